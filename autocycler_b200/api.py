"""Host-side mirror of the reference's interface for the compress path, over the C ABI of
libautocycler_gpu.so (include/autocycler_gpu.h).

Names and argument meaning follow the reference (rrwick/Autocycler v0.6.1):
    KmerGraph(k_size).add_sequences(seqs, assembly_count)      kmer_graph.rs:79-90
    UnitigGraph.from_kmer_graph(kmer_graph)                    unitig_graph.rs:36-48
    simplify_structure(unitig_graph, seqs)                     graph_simplification.rs:26-40
    unitig_graph.save_gfa(path, seqs)                          unitig_graph.rs:317-331
    load_sequences(assemblies_dir, k_size, max_contigs)        compress.rs:98-133
    compress(assemblies_dir, autocycler_dir, k_size, ...)      compress.rs:32-50

There is no CPU path here: if the CUDA library is missing, or no device is present, every entry
point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libautocycler_gpu.so")

AC_OK = 0


class AutocyclerGpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class AcConfig(C.Structure):
    _fields_ = [("k", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p), ("keep_positions", C.c_uint32),
                ("n_devices", C.c_int32), ("devices", C.POINTER(C.c_int32))]


class AcCounts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_kmers", "n_unitigs", "n_links", "total_length", "seq_bytes", "n_fwd_pos",
                                         "n_rev_pos", "n_next", "n_sequences", "n_path_steps", "length_before_simplify")]


class AcUnitigs(C.Structure):
    _fields_ = [("number", C.POINTER(C.c_uint32)), ("seq_off", C.POINTER(C.c_uint64)), ("seq", C.POINTER(C.c_uint8)),
                ("depth", C.POINTER(C.c_double)),
                ("fpos_off", C.POINTER(C.c_uint64)), ("fpos", C.POINTER(C.c_uint32)), ("fpos_id_strand", C.POINTER(C.c_uint16)),
                ("rpos_off", C.POINTER(C.c_uint64)), ("rpos", C.POINTER(C.c_uint32)), ("rpos_id_strand", C.POINTER(C.c_uint16)),
                ("next_off", C.POINTER(C.c_uint64)), ("next", C.POINTER(C.c_int32))]


class AcTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d", "pack", "insert", "adjacency", "boundaries", "runs", "unitigs", "links", "seed_sort", "emit", "d2h",
                                        "device_total", "host_graph", "host_simplify", "host_gfa", "sample", "device_simplify", "device_gfa", "insert_kernel", "reserved0")] + \
               [(n, C.c_uint64) for n in ("insert_occurrences", "table_capacity", "table_used", "kernel_launches", "h2d_bytes", "d2h_bytes")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


EXPORTS = ["ac_last_error", "ac_version", "ac_create", "ac_destroy", "ac_add_sequence", "ac_clear_sequences", "ac_upload",
           "ac_build", "ac_compress", "ac_simplify", "ac_merge_linear_paths", "ac_renumber_unitigs", "ac_load_gfa", "ac_bind_host_to_device", "ac_decompress_gfa", "ac_pairwise_distances", "ac_distance_matrix_text", "ac_sequence_reconstruct", "ac_counts_get", "ac_unitigs_copy", "ac_path_copy", "ac_gfa_size", "ac_gfa_copy",
           "ac_timings_get", "ac_compress_dir", "ac_compress_dir_devices", "ac_load_sequences", "ac_sequence_get",
           "ac_build_local", "ac_entries_count", "ac_entries_export", "ac_entries_merge", "ac_runs_local", "ac_runs_export",
           "ac_runs_import", "ac_runs_import_padded", "ac_build_finish", "ac_compress_finish", "ac_gfa_data",
           "ac_compress_finish_split", "ac_path_tokens_export", "ac_path_lines_render", "ac_path_lines_data", "ac_upload_shard", "ac_strand_block"]

_libs = {}


def load_library(path=None):
    """Loads the C-ABI library (default: the in-tree CUDA build) and declares its prototypes."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise AutocyclerGpuError(-2, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                     "(the GPU path has no CPU fallback)")
    lib = C.CDLL(path)
    lib.ac_last_error.restype = C.c_char_p
    lib.ac_last_error.argtypes = [C.c_void_p]
    lib.ac_version.restype = C.c_char_p
    lib.ac_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(AcConfig)]
    lib.ac_destroy.argtypes = [C.c_void_p]
    lib.ac_destroy.restype = None
    lib.ac_add_sequence.argtypes = [C.c_void_p, C.c_uint16, C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p]
    lib.ac_clear_sequences.argtypes = [C.c_void_p]
    for name in ("ac_upload", "ac_build", "ac_compress", "ac_simplify", "ac_renumber_unitigs"):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.ac_merge_linear_paths.argtypes = [C.c_void_p, C.c_int]
    lib.ac_load_gfa.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
    lib.ac_bind_host_to_device.argtypes = [C.c_int32]
    lib.ac_pairwise_distances.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64]
    lib.ac_distance_matrix_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.ac_decompress_gfa.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32]
    lib.ac_sequence_reconstruct.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.ac_counts_get.argtypes = [C.c_void_p, C.POINTER(AcCounts)]
    lib.ac_unitigs_copy.argtypes = [C.c_void_p, C.POINTER(AcUnitigs)]
    lib.ac_path_copy.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), C.c_uint64, C.POINTER(C.c_uint64)]
    lib.ac_gfa_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.ac_gfa_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ac_timings_get.argtypes = [C.c_void_p, C.POINTER(AcTimings)]
    lib.ac_compress_dir.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32]
    lib.ac_compress_dir_devices.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_int32, C.c_int32]
    lib.ac_load_sequences.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    lib.ac_sequence_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint16), C.POINTER(C.c_uint64), C.c_char_p, C.c_uint64,
                                    C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    lib.ac_build_local.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.ac_entries_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.ac_entries_export.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ac_entries_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ac_runs_local.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.ac_runs_export.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ac_runs_import.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ac_runs_import_padded.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    lib.ac_build_finish.argtypes = [C.c_void_p]
    lib.ac_compress_finish.argtypes = [C.c_void_p]
    lib.ac_compress_finish_split.argtypes = [C.c_void_p]
    lib.ac_upload_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.ac_strand_block.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.ac_path_tokens_export.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    lib.ac_path_lines_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ac_path_lines_data.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.ac_gfa_data.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    _libs[path] = lib
    return lib


class Sequence:
    """sequence.rs:19-28: the padded forward strand plus what save_gfa prints."""

    def __init__(self, id, forward_seq, filename, contig_header, length):
        self.id, self.forward_seq, self.filename, self.contig_header, self.length = id, forward_seq, filename, contig_header, length

    @staticmethod
    def new_with_seq(id, seq, filename, contig_header, length, half_k):   # sequence.rs:31-59
        if any(c not in "ACGT" for c in seq):
            raise AutocyclerGpuError(-6, f"{filename} contains non-ACGT characters")
        return Sequence(id, "." * half_k + seq + "." * half_k, filename, contig_header, length)


def stream_handle(stream):
    """The value of ac_config.stream for a caller's stream: None -> NULL (the library creates a private, non-blocking stream); a
    cudaStream_t -> itself; 0 — CUDA's legacy default stream, what torch.cuda.current_stream().cuda_stream is unless the caller switched
    streams — -> its explicit handle cudaStreamLegacy (0x1), because NULL already means "private"."""
    return None if stream is None else (int(stream) or 1)


class _Handle:
    def __init__(self, lib, k, device=0, stream=None, keep_positions=False, devices=None):
        self.lib = lib
        self.k = k
        self.stream = stream             # what the caller named (None: the library's own); stream_handle() is what the library is given
        self.ptr = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices) if devices else None      # several GPUs driven by this one process (ac_config.n_devices)
        cfg = AcConfig(k, device, stream_handle(stream), 1 if keep_positions else 0, len(devices) if devices else 0, devs)
        rc = lib.ac_create(C.byref(self.ptr), C.byref(cfg))
        if rc != AC_OK:
            raise AutocyclerGpuError(rc, lib.ac_last_error(None).decode())

    def check(self, rc):
        if rc != AC_OK:
            raise AutocyclerGpuError(rc, self.lib.ac_last_error(self.ptr).decode())

    def runs_on(self, cuda_stream):
        """True when the library was told to run on exactly this cudaStream_t (then work a caller enqueues there is ordered with the
        library's by the stream itself); a handle with a private stream is ordered with nobody and needs the device to settle."""
        return self.stream is not None and int(self.stream) == int(cuda_stream)

    def close(self):
        if self.ptr:
            self.lib.ac_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KmerGraph:
    """kmer_graph.rs:73-90.  The k-mers live in a hash table in HBM; `add_sequences` stages and uploads the strands,
    the table itself is built by UnitigGraph.from_kmer_graph (one fused device pipeline)."""

    def __init__(self, k_size, device=0, stream=None, lib=None, keep_positions=False, devices=None):
        self.k_size = k_size
        self._h = _Handle(lib or load_library(), k_size, device, stream, keep_positions, devices)
        self.assembly_count = 0
        self.sequences = []

    def add_sequences(self, seqs, assembly_count, upload=True):
        h = self._h
        h.check(h.lib.ac_clear_sequences(h.ptr))
        for s in seqs:
            fwd = s.forward_seq if isinstance(s.forward_seq, bytes) else s.forward_seq.encode()
            h.check(h.lib.ac_add_sequence(h.ptr, s.id, fwd, len(fwd), s.filename.encode(), s.contig_header.encode()))
        self.assembly_count = assembly_count
        self.sequences = list(seqs)
        if upload:
            self.upload()

    def upload(self):
        self._h.check(self._h.lib.ac_upload(self._h.ptr))


class UnitigGraph:
    """unitig_graph.rs:28-48."""

    def __init__(self, kmer_graph):
        self._kg = kmer_graph
        self._h = kmer_graph._h
        self.k_size = kmer_graph.k_size

    @staticmethod
    def from_kmer_graph(kmer_graph):
        g = UnitigGraph(kmer_graph)
        g._h.check(g._h.lib.ac_build(g._h.ptr))
        return g

    @staticmethod
    def compress(kmer_graph):
        """compress.rs:42-47 as one device pipeline (ac_compress): the simplified, renumbered graph; its GFA text is ready
        (gfa_view / gfa_bytes / save_gfa), the graph arrays are fetched from HBM when first asked for."""
        g = UnitigGraph(kmer_graph)
        g._h.check(g._h.lib.ac_compress(g._h.ptr))
        return g

    @staticmethod
    def from_gfa_lines(gfa_lines, lib=None, device=0):   # unitig_graph.rs:55-74 -> (UnitigGraph, [Sequence without bytes])
        text = gfa_lines if isinstance(gfa_lines, (bytes, str)) else "\n".join(l.rstrip("\n") for l in gfa_lines) + "\n"
        text = text.encode() if isinstance(text, str) else text
        kg = KmerGraph(51, device=device, lib=lib)          # the handle's k is replaced by the file's KM:i: value
        g = UnitigGraph(kg)
        g._h.check(g._h.lib.ac_load_gfa(g._h.ptr, text, len(text)))
        g.k_size = None
        seqs = []
        n = g.counts().n_sequences
        for i in range(n):
            sid, ln = C.c_uint16(), C.c_uint64()
            fn, hd = C.create_string_buffer(4096), C.create_string_buffer(65536)
            g._h.check(g._h.lib.ac_sequence_get(g._h.ptr, i, C.byref(sid), C.byref(ln), None, 0, fn, len(fn), hd, len(hd)))
            seqs.append(Sequence(sid.value, None, fn.value.decode(), hd.value.decode(), ln.value))
        return g, seqs

    @staticmethod
    def from_gfa_file(gfa_filename, lib=None, device=0):   # unitig_graph.rs:50-53
        with open(gfa_filename, "rb") as f:
            return UnitigGraph.from_gfa_lines(f.read(), lib=lib, device=device)

    def counts(self):
        c = AcCounts()
        self._h.check(self._h.lib.ac_counts_get(self._h.ptr, C.byref(c)))
        return c

    def kmer_count(self):           # KmerGraph.kmers.len(), compress.rs:152
        return self.counts().n_kmers

    def total_length(self):         # unitig_graph.rs:474-476
        return self.counts().total_length

    def link_count(self):           # unitig_graph.rs:478-507 (.1)
        return self.counts().n_links

    def timings(self):
        t = AcTimings()
        self._h.check(self._h.lib.ac_timings_get(self._h.ptr, C.byref(t)))
        return t

    def unitigs(self, positions=False):
        """-> list of dicts (number, seq, depth, forward_next, reverse_next[, forward_positions, reverse_positions])
        in the graph's current order."""
        c = self.counts()
        U = c.n_unitigs
        number = (C.c_uint32 * U)(); seq_off = (C.c_uint64 * (U + 1))(); seq = (C.c_uint8 * max(1, c.seq_bytes))()
        depth = (C.c_double * U)(); next_off = (C.c_uint64 * (2 * U + 1))(); nxt = (C.c_int32 * max(1, c.n_next))()
        u = AcUnitigs()
        u.number = number; u.seq_off = seq_off; u.seq = seq; u.depth = depth; u.next_off = next_off; u.next = nxt
        if positions:
            fo = (C.c_uint64 * (U + 1))(); fp = (C.c_uint32 * max(1, c.n_fwd_pos))(); fi = (C.c_uint16 * max(1, c.n_fwd_pos))()
            ro = (C.c_uint64 * (U + 1))(); rp = (C.c_uint32 * max(1, c.n_rev_pos))(); ri = (C.c_uint16 * max(1, c.n_rev_pos))()
            u.fpos_off = fo; u.fpos = fp; u.fpos_id_strand = fi; u.rpos_off = ro; u.rpos = rp; u.rpos_id_strand = ri
        self._h.check(self._h.lib.ac_unitigs_copy(self._h.ptr, C.byref(u)))
        raw = bytes(seq)
        out = []
        for i in range(U):
            d = {"number": number[i], "seq": raw[seq_off[i]:seq_off[i + 1]].decode(), "depth": depth[i],
                 "forward_next": [nxt[x] for x in range(next_off[2 * i], next_off[2 * i + 1])],
                 "reverse_next": [nxt[x] for x in range(next_off[2 * i + 1], next_off[2 * i + 2])]}
            if positions:
                fmt = lambda p, t: f"{t & 0x7FFF}{'+' if t & 0x8000 else '-'}{p}"   # position.rs:54-58
                d["forward_positions"] = [fmt(fp[x], fi[x]) for x in range(fo[i], fo[i + 1])]
                d["reverse_positions"] = [fmt(rp[x], ri[x]) for x in range(ro[i], ro[i + 1])]
            out.append(d)
        return out

    def get_unitig_path_for_sequence_i32(self, seq_index):   # unitig_graph.rs:467-472
        n = C.c_uint64()
        self._h.check(self._h.lib.ac_path_copy(self._h.ptr, seq_index, None, 0, C.byref(n)))
        buf = (C.c_int32 * max(1, n.value))()
        self._h.check(self._h.lib.ac_path_copy(self._h.ptr, seq_index, buf, n.value, C.byref(n)))
        return list(buf[:n.value])

    def gfa_bytes(self):            # the bytes save_gfa writes
        n = C.c_uint64()
        self._h.check(self._h.lib.ac_gfa_size(self._h.ptr, C.byref(n)))
        out = bytearray(n.value)
        if n.value:
            buf = (C.c_char * n.value).from_buffer(out)     # the library writes straight into the result, no second copy
            self._h.check(self._h.lib.ac_gfa_copy(self._h.ptr, buf, n.value))
            del buf
        return out

    def gfa_view(self):             # the same bytes without a copy: a memoryview of the library's buffer, valid until the next call on this graph
        n = C.c_uint64(); ptr = C.c_void_p()
        self._h.check(self._h.lib.ac_gfa_data(self._h.ptr, C.byref(ptr), C.byref(n)))
        return memoryview((C.c_char * n.value).from_address(ptr.value)) if n.value else memoryview(b"")

    def pairwise_contig_distances(self):   # cluster.rs:132-151 -> S x S list of lists (row a, column b)
        S = self.counts().n_sequences
        buf = (C.c_double * max(1, S * S))()
        self._h.check(self._h.lib.ac_pairwise_distances(self._h.ptr, buf, S * S))
        return [[buf[a * S + b] for b in range(S)] for a in range(S)]

    def distance_matrix_text(self):   # cluster.rs:160-176
        n = C.c_uint64()
        self._h.check(self._h.lib.ac_distance_matrix_text(self._h.ptr, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        self._h.check(self._h.lib.ac_distance_matrix_text(self._h.ptr, buf, n.value, C.byref(n)))
        return buf.raw[:n.value].decode()

    def renumber_unitigs(self):   # unitig_graph.rs:295-315
        self._h.check(self._h.lib.ac_renumber_unitigs(self._h.ptr))

    def reconstruct_original_sequence(self, index):   # unitig_graph.rs:383-388, by position in the sequence list
        n = C.c_uint64()
        self._h.check(self._h.lib.ac_sequence_reconstruct(self._h.ptr, index, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        self._h.check(self._h.lib.ac_sequence_reconstruct(self._h.ptr, index, buf, n.value, C.byref(n)))
        return buf.raw[:n.value].decode()

    def save_gfa(self, gfa_filename, sequences=None, use_other_colour=False):   # unitig_graph.rs:317-331
        with open(gfa_filename, "wb") as f:
            f.write(self.gfa_bytes())


def simplify_structure(graph, seqs=None):   # graph_simplification.rs:26-40
    graph._h.check(graph._h.lib.ac_simplify(graph._h.ptr))


def decompress(in_gfa, out_dir=None, out_file=None, lib=None, device=0):   # decompress.rs:27-38
    lib = lib or load_library()
    rc = lib.ac_decompress_gfa(os.fsencode(in_gfa), os.fsencode(out_dir) if out_dir else None, os.fsencode(out_file) if out_file else None, device, 0)
    if rc != AC_OK:
        raise AutocyclerGpuError(rc, lib.ac_last_error(None).decode())


def merge_linear_paths(graph, seqs=()):   # graph_simplification.rs:315-371; seqs=None/[] merges without regard to the paths
    graph._h.check(graph._h.lib.ac_merge_linear_paths(graph._h.ptr, 1 if seqs is not None and len(seqs) else 0))


def load_sequences(assemblies_dir, k_size, max_contigs=25, threads=8, lib=None, device=0, stream=None):
    """compress.rs:98-133 -> (KmerGraph holding the staged sequences, [Sequence], assembly_count)."""
    kg = KmerGraph(k_size, device=device, lib=lib, stream=stream)
    h = kg._h
    count = C.c_uint64()
    h.check(h.lib.ac_load_sequences(h.ptr, os.fsencode(assemblies_dir), max_contigs, threads, C.byref(count)))
    seqs = []
    i = 0
    while True:
        sid = C.c_uint16(); length = C.c_uint64()
        if h.lib.ac_sequence_get(h.ptr, i, C.byref(sid), C.byref(length), None, 0, None, 0, None, 0) != AC_OK:
            break
        fwd = C.create_string_buffer(length.value + k_size); fn = C.create_string_buffer(4096); hd = C.create_string_buffer(1 << 16)
        h.check(h.lib.ac_sequence_get(h.ptr, i, None, None, fwd, len(fwd), fn, len(fn), hd, len(hd)))
        seqs.append(Sequence(sid.value, fwd.value.decode(), fn.value.decode(), hd.value.decode(), length.value))
        i += 1
    kg.sequences = seqs
    kg.assembly_count = count.value
    return kg, seqs, count.value


def compress(assemblies_dir, autocycler_dir, k_size=51, max_contigs=25, threads=8, device=0, verbose=False, lib=None, devices=None):
    """compress.rs:32-50: writes <autocycler_dir>/input_assemblies.gfa and .yaml.  devices=[...]: sharded by file over several GPUs."""
    lib = lib or load_library()
    devs = list(devices) if devices else [device]
    rc = lib.ac_compress_dir_devices(os.fsencode(assemblies_dir), os.fsencode(autocycler_dir), k_size, max_contigs, threads,
                                     (C.c_int32 * len(devs))(*devs), len(devs), 1 if verbose else 0)
    if rc != AC_OK:
        raise AutocyclerGpuError(rc, lib.ac_last_error(None).decode())
