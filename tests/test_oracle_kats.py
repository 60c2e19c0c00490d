"""Pins the CPU oracle to every known-answer test the reference holds for the compress path.
Each test names the reference test it restates (file:line under /root/reference/src)."""
import gzip
import os
import random

import pytest

import oracle_lib as o

SEQ20 = "ACGACTGACATCAGCACTGC"


def gfa(golden_dir, n):
    return open(os.path.join(golden_dir, f"ref_test_gfa_{n}.gfa")).read()


def test_position_display():  # position.rs:64-72
    assert o.position_display(1, True, 123) == "1+123"
    assert o.position_display(2, False, 456) == "2-456"
    assert o.position_display(32767, True, 4294967295) == "32767+4294967295"


def test_reverse_complement():  # misc.rs:324-342
    assert o.reverse_complement("ACGT..") == "..ACGT"
    assert o.reverse_complement("AXC") == "GNT"


def test_kmer_graph_count_and_iterate():  # kmer_graph.rs:199-212, 266-282
    expected = ["..ACG", "..GCA", ".ACGA", ".GCAG", "ACATC", "ACGAC", "ACTGA", "ACTGC", "AGCAC", "AGTCG",
                "AGTGC", "ATCAG", "ATGTC", "CACTG", "CAGCA", "CAGTC", "CAGTG", "CATCA", "CGACT", "CGT..",
                "CTGAC", "CTGAT", "CTGC.", "GACAT", "GACTG", "GATGT", "GCACT", "GCAGT", "GCTGA", "GTCAG",
                "GTCGT", "GTGCT", "TCAGC", "TCAGT", "TCGT.", "TGACA", "TGATG", "TGC..", "TGCTG", "TGTCA"]
    got = o.kmers_sorted(SEQ20, 5)
    assert len(got) == 40
    assert [g.split(":")[0] for g in got] == expected
    # first forward k-mer sits at 1+0, and the k-mer display format is kmer_graph.rs:189-197
    assert "..ACG:1+0" in got


def test_next_kmers():  # kmer_graph.rs:214-238
    assert o.neighbour_kmers(SEQ20, 5, "ACATC", True) == ["CATCA"]
    assert o.neighbour_kmers(SEQ20, 5, "CACTG", True) == ["ACTGA", "ACTGC"]
    assert o.neighbour_kmers(SEQ20, 5, "ACTGA", True) == ["CTGAC", "CTGAT"]
    assert o.neighbour_kmers(SEQ20, 5, "AAAAA", True) == []


def test_prev_kmers():  # kmer_graph.rs:240-264
    assert o.neighbour_kmers(SEQ20, 5, "CATCA", False) == ["ACATC"]
    assert o.neighbour_kmers(SEQ20, 5, "CTGAC", False) == ["ACTGA", "GCTGA"]
    assert o.neighbour_kmers(SEQ20, 5, "ACTGC", False) == ["CACTG", "GACTG"]
    assert o.neighbour_kmers(SEQ20, 5, "AAAAA", False) == []


def test_unitig_from_kmers():  # unitig.rs:410-440
    assert o.unitig_from_kmers("ACGCATAGCACTAGCTACGA", 5, 4) == ["GCATAGC", "GCTATGC", "ATA", "TAT"]


def test_unitig_shift_ops():  # unitig.rs:457-555
    seg = "S\t1\tGCTGAAGGGC\tDP:f:1"
    assert o.unitig_shift(seg, 0, 2) == ["TGAAGGGC", "GCCCTTCA", "102,202", "890,790"]
    assert o.unitig_shift(seg, 1, 2) == ["GCTGAAGG", "CCTTCAGC", "100,200", "892,792"]
    assert o.unitig_shift(seg, 2, seq="AC") == ["ACGCTGAAGGGC", "GCCCTTCAGCGT", "98,198", "890,790"]
    assert o.unitig_shift(seg, 3, seq="AC") == ["GCTGAAGGGCAC", "GTGCCCTTCAGC", "100,200", "888,788"]


def test_find_best_match_1():  # compress.rs:281-314
    f = o.find_best_match
    assert f(["...ACGT"]) == "...ACGT"
    assert f(["...ACGT", "..GACGT"]) == "..GACGT"
    assert f(["..GACGT", "...ACGT"]) == "..GACGT"
    assert f(["...GAAA", "...CAAA", "...TAAA"]) == "...CAAA"
    assert f(["...ACGT", "..GACGT", "..CACGT", "..GACGT", "..CACGT"]) == "..CACGT"
    assert f(["...ACGT", "..GACGT", "..GACGT", ".AGACGT", ".CGACGT"]) == ".AGACGT"
    assert f(["...ACGT", ".CGACGT", "..GACGT", ".AGACGT", ".CGACGT"]) == ".CGACGT"


def test_find_best_match_2():  # compress.rs:316-344
    f = o.find_best_match
    assert f(["ACGT..."]) == "ACGT..."
    assert f(["ACGT...", "ACGTT.."]) == "ACGTT.."
    assert f(["..GACGT", "...ACGT"]) == "..GACGT"
    assert f(["GAAA...", "CAAA...", "TAAA..."]) == "CAAA..."
    assert f(["CACG...", "GACGT..", "CACGT..", "GACGT..", "CACGT.."]) == "CACGT.."
    assert f(["AGAC...", "AGACG..", "AGACG..", "AGACGT.", "CGACGT."]) == "AGACGT."


def test_load_sequences_counts(tmp_path):  # compress.rs:346-356
    (tmp_path / "a.fasta").write_text(">a1\nACGT\n")
    (tmp_path / "b.fasta").write_text(">b1\nACGT\n>b2\nACGT\n")
    (tmp_path / "c.fasta").write_text(">c1\nACGT\n>c2\nACGT\n>c3\nACGT\n")
    count, seqs = o.load_sequences(str(tmp_path), 3)
    assert count == 3 and len(seqs) == 6


def test_load_sequences_duplicate_name(tmp_path):  # compress.rs:358-369
    (tmp_path / "a.fasta").write_text(">a1\nACGT\n")
    (tmp_path / "c.fasta").write_text(">c1\nACGT\n>c1\nACGT\n>c3\nACGT\n")
    with pytest.raises(o.OracleError, match="duplicate name"):
        o.load_sequences(str(tmp_path), 3)


def test_whitespace_and_padding(tmp_path):  # tests.rs:170-188
    (tmp_path / "assembly.fasta").write_text(">name abc  def\tghi\nCTTATGAGCAGTCCTTAACGTAGCGGT\n")
    count, seqs = o.load_sequences(str(tmp_path), 11)
    assert count == 1
    sid, fn, hd, length, fwd = seqs[0]
    assert fn == "assembly.fasta" and hd == "name abc def ghi"
    assert fwd == ".....CTTATGAGCAGTCCTTAACGTAGCGGT....."


def test_common_start_seq(golden_dir):  # graph_simplification.rs:540-559
    g = "H\tVN:Z:1.0\tKM:i:9\nS\t1\tACGATCAGC\tDP:f:1\nS\t2\tACTATCAGC\tDP:f:1\nS\t3\tACTACGACT\tDP:f:1\n"
    assert o.gfa_common_seq(g, "1+,2+,3+", False) == "AC"
    assert o.gfa_common_seq(g, "1+,2+,3-", False) == "A"
    assert o.gfa_common_seq(g, "1+,2-,3-", False) == ""


def test_common_end_seq(golden_dir):  # graph_simplification.rs:561-580
    g = "H\tVN:Z:1.0\tKM:i:9\nS\t1\tACGATCAGC\tDP:f:1\nS\t2\tACTATCAGC\tDP:f:1\nS\t3\tACTACGACT\tDP:f:1\n"
    assert o.gfa_common_seq(g, "1+,2+,3+", True) == ""
    assert o.gfa_common_seq(g, "1-,2-,3+", True) == "T"
    assert o.gfa_common_seq(g, "1-,2-,3-", True) == "GT"


def test_exclusive_inputs_and_outputs(golden_dir):  # graph_simplification.rs:582-625
    ex = o.gfa_exclusive(gfa(golden_dir, 1))
    assert ex[1] == ("2+,3-", "")
    assert ex[2] == ("", "") and ex[3] == ("", "")
    assert ex[4] == ("", "7-,8+")
    assert ex[5] == ("", "") and ex[6] == ("", "")
    assert ex[7] == ("9-,9+", "")
    assert ex[8] == ("", "10-")
    assert ex[9] == ("", "")
    assert ex[10] == ("", "8-")


def test_simplify_structure_1(golden_dir):  # graph_simplification.rs:627-654
    before = [s for _, s in o.gfa_unitig_seqs(gfa(golden_dir, 1))]
    assert before == ["TTCGCTGCGCTCGCTTCGCTTT", "TGCCGTCGTCGCTGTGCA", "TGCCTGAATCGCCTA", "GCTCGGCTCG", "CGAACCAT",
                      "TACTTGT", "GCCTT", "ATCT", "GC", "T"]
    after = [s for _, s in o.gfa_unitig_seqs(gfa(golden_dir, 1), simplify=True)]
    assert after == ["GCATTCGCTGCGCTCGCTTCGCTTT", "TGCCGTCGTCGCTGT", "CTGAATCGCCTA", "GCTCGGCTCGA", "CGAACCAT",
                     "TACTTGT", "GCCT", "TCT", "GC", "T"]


def test_simplify_structure_2(golden_dir):  # graph_simplification.rs:656-671
    before = [s for _, s in o.gfa_unitig_seqs(gfa(golden_dir, 2))]
    assert before == ["ACCGCTGCGCTCGCTTCGCTCT", "ATGAT", "GCGC"]
    after = [s for _, s in o.gfa_unitig_seqs(gfa(golden_dir, 2), simplify=True)]
    assert after == ["CACCGCTGCGCTCGCTTCGCTCTAT", "CG", "G"]


def _merged(golden_dir, n):
    text = o.gfa_merge_linear_paths(gfa(golden_dir, n))
    segs = {int(l.split("\t")[1]): l.split("\t")[2] for l in text.splitlines() if l[0] == "S"}
    links = sorted(tuple(l.split("\t")[1:5]) for l in text.splitlines() if l[0] == "L")
    return segs, links


def test_can_merge(golden_dir):  # graph_simplification.rs:701-740
    starts, ends = o.gfa_merge_fixed_sets(gfa(golden_dir, 14))
    assert starts == [5, 8, 12, 19, 22]
    assert ends == [8, 17, 19, 22, 37]


def test_merge_linear_paths_1(golden_dir):  # graph_simplification.rs:742-764
    segs, links = _merged(golden_dir, 3)
    assert segs == {8: "TTCGCTGCGCTCGCTTCGCTTTTGCACAGCGACGACGGCATGCCTGAATCGCCTA", 9: "GCTCGGCTCGATGGTTCG", 10: "TACTTGTAAGGC"}
    assert links == sorted([("8", "+", "9", "+"), ("9", "-", "8", "-"), ("9", "+", "9", "-"), ("8", "+", "10", "+"),
                            ("10", "-", "8", "-"), ("10", "+", "10", "+"), ("10", "-", "10", "-")])


def test_merge_linear_paths_2(golden_dir):  # graph_simplification.rs:766-783
    segs, links = _merged(golden_dir, 4)
    assert segs == {6: "ACGACTACGAGCACGAGTCGTCGTCGTAACTGACT", 7: "GCTCGGTG"}
    assert links == sorted([("6", "+", "6", "+"), ("6", "-", "6", "-"), ("7", "+", "7", "+"), ("7", "-", "7", "-")])


def test_merge_linear_paths_3(golden_dir):  # graph_simplification.rs:785-793
    segs, _ = _merged(golden_dir, 5)
    assert len(segs) == 5 and segs[7] == "AAATGCGACTGTG"


def test_merge_linear_paths_4(golden_dir):  # graph_simplification.rs:795-801
    assert len(_merged(golden_dir, 14)[0]) == 11


@pytest.mark.parametrize("n", range(1, 15))
def test_reference_gfa_fixtures_load_and_check_links(golden_dir, n):  # test_gfa.rs:15-287 via from_gfa_lines
    text = gfa(golden_dir, n)
    again = o.gfa_roundtrip(text)
    # S/L/P line multiset is preserved (depth is re-rendered with two decimals, as save_gfa does)
    assert sum(1 for l in again.splitlines() if l[0] == "S") == sum(1 for l in text.splitlines() if l[0] == "S")
    assert sorted(l for l in again.splitlines() if l[0] == "L") == sorted(l for l in text.splitlines() if l[0] == "L")
    assert o.gfa_roundtrip(again) == again


def test_gfa_14_roundtrip_is_identity(golden_dir):  # test_gfa.rs:238-287 is itself a save_gfa-format file
    text = gfa(golden_dir, 14)
    assert o.gfa_roundtrip(text) == text


# ---- the reference's end-to-end invariants, tests.rs:75-167 -------------------------------------

def high_level(tmp_path, seqs, k):
    """tests.rs:75-128 test_high_level"""
    asm = tmp_path / f"asm_k{k}"; rec = tmp_path / f"rec_k{k}"
    asm.mkdir(); rec.mkdir()
    a, b, c, d, e = seqs
    (asm / "a.fasta").write_text(a)
    (asm / "b.fna").write_text(b)
    (asm / "c.fa").write_text(c)
    (asm / "d.fasta.gz").write_bytes(gzip.compress(d.encode()))
    (asm / "e.fna.gz").write_bytes(gzip.compress(e.encode()))
    (asm / "e.xyz").write_text(a)  # bad extension, not included
    count, loaded = o.load_sequences(str(asm), k)
    assert count == 5
    gfa_1, yaml, st = o.compress_dir(str(asm), k)
    gfa_2 = o.gfa_roundtrip(gfa_1)
    assert gfa_1 == gfa_2                                    # tests.rs:108-112
    o.decompress(gfa_1, str(rec))                            # tests.rs:114-127
    assert (rec / "a.fasta").read_text() == a
    assert (rec / "b.fna").read_text() == b
    assert (rec / "c.fa").read_text() == c
    assert gzip.decompress((rec / "d.fasta.gz").read_bytes()).decode() == d
    assert gzip.decompress((rec / "e.fna.gz").read_bytes()).decode() == e
    return gfa_1


FIXED = [">a\nCTTATGAGCAGTCCTTAACGTAGCGGTGTGTGGCTTTGAGAAGTTAGCGGTGGCGAGCTACATCCTGGCTCCAAT\n",
         ">b\nACCGTTACGTTAAGGACTGCTCATAAGATTGGAGCCAGGATGTAGCTCGCCACGGCTAACTTCTCAAAGCGGCAC\n",
         ">c\nCATCCTGGCTCCAATCTTATGAGCAGTCCTTAACGTAACGGTGTGTGGCTTTGAGAAGTTAGCCGTGGCGAGATA\n",
         ">d\nGGACTGCTCATAAGATTGGAGCCAGGATGTAGCTCGCCACGGCTAACTTCTCAAAGCCACACACCGTTACGTTAA\n",
         ">e\nTTGAGAAGTTAGCCGTGGCGAGCTACATCCTGGCTCCAATCTTATGAGCAGTCCTTAACGTAACGGTGTGTGGCC\n"]


@pytest.mark.parametrize("k", [1, 5, 9, 13, 51])
def test_fixed_seqs(tmp_path, k):  # tests.rs:131-148
    high_level(tmp_path, FIXED, k)


@pytest.mark.parametrize("length", [10, 20, 50, 100])
def test_random_seqs(tmp_path, length):  # tests.rs:151-167 (the reference seeds Rust's StdRng; any seeded iid ACGT will do)
    for seed in [0, 5, 10, 15, 20]:
        seqs = []
        for j, name in enumerate("abcde"):
            rng = random.Random(seed + j)
            seqs.append(f">{name}\n" + "".join(rng.choice("ACGT") for _ in range(length)) + "\n")
        for k in [3, 5, 7, 9]:
            sub = tmp_path / f"s{seed}k{k}"; sub.mkdir()
            high_level(sub, seqs, k)
