"""Shared input generators for the parity tests (oracle vs CUDA path, and oracle vs the host-emulated
device logic).  Every case is a list of files: [(filename, [(header, sequence)])]."""
import random


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, rate):
    out = []
    for c in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(rng.choice("ACGT"))
            out.append(c)
        elif r < rate:
            out.append(rng.choice([b for b in "ACGT" if b != c]))
        else:
            out.append(c)
    return "".join(out)


def random_case(seed, k):
    """Adversarial little genomes: rotations, strand flips, shared ends, inverted and tandem repeats,
    homopolymers, low-complexity alphabets, multi-contig files, linear contigs whose ends match nothing."""
    rng = random.Random(seed)
    style = rng.randrange(8)
    n_files = rng.randint(1, 5)
    base_len = rng.randint(k + 1, max(k + 2, rng.choice([30, 60, 150, 400])))
    alphabet = "ACGT" if style != 5 else rng.choice(["AC", "AT", "ACG"])
    genome = rand_seq(rng, base_len, alphabet)
    if style == 1:   # inverted repeat
        unit = rand_seq(rng, rng.randint(k, 2 * k), alphabet)
        genome = genome[:len(genome) // 2] + unit + rand_seq(rng, rng.randint(0, 5)) + rc(unit) + genome[len(genome) // 2:]
    if style == 2:   # tandem repeat
        unit = rand_seq(rng, rng.randint(1, k), alphabet)
        genome = genome[:len(genome) // 3] + unit * rng.randint(2, 3 + (2 * k) // len(unit)) + genome[len(genome) // 3:]
    if style == 3:   # homopolymers
        genome = genome[:len(genome) // 2] + rng.choice("ACGT") * rng.randint(k - 1, 2 * k + 3) + genome[len(genome) // 2:]
    if style == 4:   # dispersed repeat
        unit = rand_seq(rng, rng.randint(k, 3 * k), alphabet)
        genome = unit + genome[:len(genome) // 2] + unit + genome[len(genome) // 2:] + (rc(unit) if rng.random() < 0.5 else unit)
    files = []
    for f in range(n_files):
        recs = []
        n_contigs = 1 if rng.random() < 0.7 else rng.randint(2, 3)
        for c in range(n_contigs):
            s = genome
            if rng.random() < 0.7:
                r = rng.randrange(len(s))
                s = s[r:] + s[:r]
            if rng.random() < 0.5:
                s = rc(s)
            if rng.random() < 0.8:
                s = mutate(rng, s, rng.choice([0.0, 0.01, 0.05]))
            if rng.random() < 0.3:
                s = s + s[:rng.randint(1, min(len(s), 3 * k))]          # circular overlap
            if rng.random() < 0.3 and len(s) > 2 * k + 4:
                a = rng.randrange(len(s) - k - 1)
                s = s[a:a + rng.randint(k, len(s) - a)]                  # linear fragment
            if style == 6 and rng.random() < 0.5:
                s = rand_seq(rng, rng.randint(k, 3 * k))                # unrelated contig: its ends stay dotted
            if style == 7:
                s = s[:k + rng.randint(0, 3)]                           # contigs barely longer than k
            if len(s) < k:
                s = s + rand_seq(rng, k - len(s))
            recs.append((f"c{c + 1} len={len(s)}", s))
        files.append((f"asm_{f:02d}.fasta", recs))
    return files


def write_case(files, directory):
    import os
    os.makedirs(directory, exist_ok=True)
    for fn, recs in files:
        with open(os.path.join(directory, fn), "w") as f:
            for header, seq in recs:
                f.write(f">{header}\n{seq}\n")
