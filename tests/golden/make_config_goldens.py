"""Generates tests/golden/config_goldens.json: SHA-256 of the ORACLE's input_assemblies.gfa for the
BASELINE.json configs that are too large to compare on every run (the oracle needs ~100 s for cfg2).
Run in the build container:  python tests/golden/make_config_goldens.py cfg2 [cfg3:51 cfg5:51:16 ...]
A spec is name[:k[:n_assemblies]]; with n_assemblies the first n assemblies of the config are taken (the weak-scaling
inputs of bench.py --gpus N are the first 8N assemblies of cfg5) and the key carries an _n<count> suffix.
The synthetic inputs are regenerated deterministically (autocycler_b200/synth.py), so only hashes are stored."""
import hashlib
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as o  # noqa: E402
from autocycler_b200 import synth  # noqa: E402

path = os.path.join(HERE, "config_goldens.json")
goldens = json.load(open(path)) if os.path.exists(path) else {}
for spec in sys.argv[1:]:
    name, _, rest = spec.partition(":")
    kk, _, nn = rest.partition(":")
    k = int(kk or 51)
    n_asm = int(nn) if nn else None
    key = f"{name}_k{k}" + (f"_n{n_asm}" if n_asm else "")
    with tempfile.TemporaryDirectory() as d:
        a = synth.make_assemblies(name, n_assemblies=n_asm)
        synth.write_assemblies(a, d)
        t = time.time()
        gfa, yaml, st = o.compress_dir(d, k, threads=8)
        dt = time.time() - t
    entry = dict(sha256=hashlib.sha256(gfa.encode()).hexdigest(), gfa_bytes=len(gfa), n_kmers=st.n_kmers,
                                   unitigs_before=st.unitigs_before, links_before=st.links_before,
                                   unitigs_after=st.unitigs_after, links_after=st.links_after,
                                   input_bases=synth.total_bases(a), oracle_seconds=round(dt, 1),
                                   oracle_stage_seconds=dict(load=round(st.t_load, 2), kmer_graph=round(st.t_kmer_graph, 2),
                                                             unitig_graph=round(st.t_unitig_graph, 2),
                                                             simplify=round(st.t_simplify, 2), gfa=round(st.t_gfa, 2)))
    print(spec, entry, flush=True)
    goldens = json.load(open(path)) if os.path.exists(path) else {}      # several generators may run side by side
    goldens[key] = entry
    json.dump(goldens, open(path, "w"), indent=1, sort_keys=True)
