"""End-to-end wall time of the `autocycler compress` CLI on cfg2: FASTA files on disk -> GFA + YAML on disk."""
import hashlib, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autocycler_b200 import synth

d = "/tmp/cfg2_in"
asm = synth.make_assemblies("cfg2")
synth.write_assemblies(asm, d)
print("bases", synth.total_bases(asm), flush=True)
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "autocycler_b200", "bin", "autocycler")
for rep in range(3):
    out = "/tmp/cfg2_out%d" % rep
    t0 = time.time()
    try:
        r = subprocess.run([exe, "compress", "-i", d, "-a", out], capture_output=True, text=True, timeout=100)
    except subprocess.TimeoutExpired:
        print("rep", rep, "TIMEOUT", flush=True); continue
    wall = time.time() - t0
    print("rep", rep, "rc", r.returncode, "wall_s", round(wall, 3), "Mbp/s", round(synth.total_bases(asm) / wall / 1e6, 1), flush=True)
    if rep == 2: print(r.stderr[-2500:])
    print("sha256", hashlib.sha256(open(out + "/input_assemblies.gfa", "rb").read()).hexdigest(), flush=True)
