"""torchrun script (needs >= 2 GPUs): the N-rank sharded build gives the same GFA as the single-GPU build.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py"""
import hashlib
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from autocycler_b200 import api, dist as acdist, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, n_asm, lens, k in [("a", 6, [300_000, 9_000], 51), ("b", 4, [120_000], 31), ("c", 5, [80_000], 91), ("cfg2", 8, None, 51)]:
    assemblies = synth.make_assemblies(name, n_assemblies=n_asm, replicon_lengths=lens, seed=None if name == "cfg2" else 1234)
    with tempfile.TemporaryDirectory() as d:
        synth.write_assemblies(assemblies, d)
        kg, seqs, count = api.load_sequences(d, k, device=local)
    kg.upload()
    lo, hi = acdist.shard_bounds(len(seqs), rank, world)
    g = acdist.from_kmer_graph_distributed(kg, lo, hi, torch.device("cuda", local))
    if rank == 0:
        api.simplify_structure(g)
        multi = hashlib.sha256(g.gfa_bytes()).hexdigest()
        kg.upload()
        g1 = api.UnitigGraph.from_kmer_graph(kg)
        api.simplify_structure(g1)
        single = hashlib.sha256(g1.gfa_bytes()).hexdigest()
        print(name, k, "world", world, "same" if multi == single else "DIFFERENT", multi[:16], flush=True)
        ok = ok and multi == single
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "OK" if ok else "FAIL", flush=True)
dist.destroy_process_group()
