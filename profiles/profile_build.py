"""One warm-up graph build, then one build between cudaProfilerStart/Stop: the target of the ncu captures
(`ncu --profile-from-start off ...`).  usage: python profiles/profile_build.py [workload] [k]"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from autocycler_b200 import api, synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 51
with tempfile.TemporaryDirectory() as d:
    synth.write_assemblies(synth.make_assemblies(workload), d)
    kg, seqs, count = api.load_sequences(d, k)


def build():
    kg.upload()
    g = api.UnitigGraph.compress(kg)
    return len(g.gfa_view())


build()
torch.cuda.synchronize()
torch.cuda.profiler.start()
n = build()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("gfa bytes", n)
