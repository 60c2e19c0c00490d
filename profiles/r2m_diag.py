"""r2m diagnostic (2 GPUs, one process, no torch/NCCL): `autocycler compress --devices 0,1` against the oracle on the multi_gpu_check cases,
repeated, with the first differing line reported; then the same under AC_SYNC_LAUNCHES and compute-sanitizer (memcheck, initcheck, racecheck)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from autocycler_b200 import synth  # noqa: E402

EXE = os.path.join(ROOT, "autocycler_b200", "bin", "autocycler")
cases = [("d", 3, [50_000], 51), ("a", 6, [300_000, 9_000], 51), ("b", 4, [120_000], 31)]
tmp = tempfile.mkdtemp()


def first_diff(a, b):
    la, lb = a.split("\n"), b.split("\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            j = next((c for c in range(min(len(x), len(y))) if x[c] != y[c]), min(len(x), len(y)))
            return f"line {i} of {len(la)}/{len(lb)} col {j}: got {x[max(0, j - 30):j + 50]!r} want {y[max(0, j - 30):j + 50]!r} (line starts {x[:12]!r})"
    return f"lengths {len(la)} vs {len(lb)} lines"


def run(d, k, devices, env=None, prefix=(), tag=""):
    out = tempfile.mkdtemp(dir=tmp)
    e = dict(os.environ); e.update(env or {})
    r = subprocess.run(list(prefix) + [EXE, "compress", "-i", d, "-a", out, "--kmer", str(k), "--devices", devices], capture_output=True, text=True, env=e, timeout=900)
    p = os.path.join(out, "input_assemblies.gfa")
    return r, (open(p).read() if os.path.exists(p) else None)


for name, n_asm, lens, k in cases:
    d = os.path.join(tmp, name)
    synth.write_assemblies(synth.make_assemblies(name, n_assemblies=n_asm, replicon_lengths=lens, seed=1234), d)
    want = oracle_lib.compress_dir(d, k)[0]
    for devices in ("0", "0,1", "0,1", "0,1"):
        r, got = run(d, k, devices)
        ok = got == want
        print(f"case {name} k={k} devices={devices}: rc={r.returncode} {'OK' if ok else 'DIFFERENT ' + (first_diff(got, want) if got else r.stderr[-400:])}", flush=True)
    r, got = run(d, k, "0,1", env={"AC_SYNC_LAUNCHES": "1"})
    print(f"case {name} sync-launches: rc={r.returncode} {'OK' if got == want else 'DIFFERENT'} {r.stderr[-300:] if r.returncode else ''}", flush=True)

name, n_asm, lens, k = cases[0]
d = os.path.join(tmp, name)
for tool in ("memcheck", "initcheck", "racecheck"):
    r, got = run(d, k, "0,1", prefix=("compute-sanitizer", "--tool", tool, "--print-limit", "20"))
    lines = [x for x in (r.stdout + r.stderr).split("\n") if "=========" in x]
    print(f"--- compute-sanitizer {tool}: rc={r.returncode} {len(lines)} lines", flush=True)
    print("\n".join(lines[:70]), flush=True)
