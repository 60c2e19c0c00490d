"""torchrun script (needs >= 2 GPUs): the N-rank sharded build gives the oracle's GFA — byte for byte on inputs the oracle finishes in
seconds, through the committed SHA-256 on cfg2 — and the same bytes as the one-GPU build.  Run by tests/test_multi_gpu.py, or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import torch
import torch.distributed as dist

from autocycler_b200 import api, dist as acdist, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
goldens = json.load(open(os.path.join(HERE, "golden", "config_goldens.json")))
quick = os.environ.get("AC_MULTI_CHECK_QUICK") is not None
cases = [("a", 6, [300_000, 9_000], 51), ("b", 4, [120_000], 31), ("c", 5, [80_000], 91), ("d", 3, [50_000], 51)]
if not quick:
    cases.append(("cfg2", 8, None, 51))
if os.environ.get("AC_MULTI_CHECK_CASES"):          # debugging: a subset by name
    cases = [c for c in cases if c[0] in os.environ["AC_MULTI_CHECK_CASES"].split(",")]


def first_diff(got, want):
    """Where two GFA texts part (for the log of a failing run)."""
    a, b = got.split(b"\n"), want.split(b"\n")
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            j = next((c for c in range(min(len(x), len(y))) if x[c] != y[c]), min(len(x), len(y)))
            return f"line {i} ({len(a)} vs {len(b)} lines, {len(got)} vs {len(want)} bytes) col {j}: got {x[max(0, j - 20):j + 40]!r} want {y[max(0, j - 20):j + 40]!r}"
    return f"{len(a)} vs {len(b)} lines, {len(got)} vs {len(want)} bytes, common part equal"


ok = True
side_stream = torch.cuda.Stream(device=local)
for case_index, (name, n_asm, lens, k) in enumerate(cases):
    assemblies = synth.make_assemblies(name, n_assemblies=n_asm, replicon_lengths=lens, seed=None if name == "cfg2" else 1234)
    with tempfile.TemporaryDirectory() as d:
        synth.write_assemblies(assemblies, d)
        # the handle runs on a stream of its own (the collectives are then fenced by device-wide synchronisation), on torch's default
        # stream, or on a torch side stream (both ordered with NCCL by the stream itself: what bench.py times)
        mode = case_index % 3
        if mode == 2:
            torch.cuda.set_stream(side_stream)
        else:
            torch.cuda.set_stream(torch.cuda.default_stream(local))
        kg, seqs, count = api.load_sequences(d, k, device=local, stream=None if mode == 0 else torch.cuda.current_stream().cuda_stream)
        expected = expected_text = None
        if rank == 0:
            if name == "cfg2":
                expected = goldens["cfg2_k51"]["sha256"]
            else:
                import oracle_lib
                expected_text = oracle_lib.compress_dir(d, k)[0].encode()
                expected = hashlib.sha256(expected_text).hexdigest()
    kg.upload()
    # contiguous blocks of FILES (assemblies shard by sorted file order); a rank may end up with none
    first_of = {}
    for i, sq in enumerate(seqs):
        first_of.setdefault(sq.filename, i)
    bounds = [first_of[f] for f in sorted(first_of, key=first_of.get)] + [len(seqs)]
    flo, fhi = acdist.shard_bounds(len(bounds) - 1, rank, world)
    g = acdist.from_kmer_graph_distributed(kg, bounds[flo], bounds[fhi], torch.device("cuda", local))
    if rank == 0:
        api.simplify_structure(g)
        multi = hashlib.sha256(g.gfa_bytes()).hexdigest()
        kg.upload()
        g1 = api.UnitigGraph.from_kmer_graph(kg)
        api.simplify_structure(g1)
        single = hashlib.sha256(g1.gfa_bytes()).hexdigest()
        good = multi == single == expected
        print(name, k, "world", world, ["private stream", "default stream", "side stream"][mode], "same as the oracle" if good else f"DIFFERENT multi={multi[:12]} single={single[:12]} oracle={expected[:12]}", flush=True)
        ok = ok and good
    dist.barrier()
    # the fused forms: rank 0 prints the whole file / every rank prints the P lines of its own sequences (what bench.py times at N > 1)
    kg.upload()
    gf = acdist.compress_distributed(kg, bounds[flo], bounds[fhi], torch.device("cuda", local))
    fused_text = bytes(gf.gfa_view()) if rank == 0 else None
    fused = hashlib.sha256(fused_text).hexdigest() if rank == 0 else None
    kg.upload()
    gs, lines = acdist.compress_distributed_split(kg, bounds[flo], bounds[fhi], torch.device("cuda", local))
    parts = [None] * world
    dist.all_gather_object(parts, bytes(lines.view()))
    if rank == 0:
        split_text = bytes(gs.gfa_view()) + b"".join(parts)
        split = hashlib.sha256(split_text).hexdigest()
        good = fused == split == expected
        print(name, k, "world", world, "fused and split-path forms", "same as the oracle" if good else f"DIFFERENT fused={fused[:12]} split={split[:12]} oracle={expected[:12]}", flush=True)
        if not good and expected_text is not None:
            for label, text in (("fused", fused_text), ("split", split_text)):
                if text != expected_text:
                    print("  ", label, first_diff(text, expected_text), flush=True)
        ok = ok and good
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "OK" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
