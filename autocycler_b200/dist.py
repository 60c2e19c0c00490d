"""Multi-GPU form of UnitigGraph.from_kmer_graph / UnitigGraph.compress: one process per GPU under torch.distributed.

Every rank stages ALL sequences on the host (1 byte per base) and owns a contiguous block of them (sorted file order).
  0. upload_sharded: every rank copies only its own strands over its PCIe link, one broadcast per block brings the others' over NVLink
  1. local k-mer table over the owned sequences                       (parallel: the dominant insert work is split N ways)
  2. ONE data-path collective for the table: all-gather of the deduplicated local tables (16 B per distinct k-mer),
     merged into every rank's table (counts add, first/last flags OR, the slot keeps the smallest occurrence, which
     names the k-mer identically on every rank)
  3. adjacency on the now global table (replicated), unitig occurrences along the owned sequences (parallel)
  4. gather of the occurrences (16 B each) to rank 0, which builds unitigs / links / seed order over all of them and
     continues exactly like the single-GPU path: host graph (from_kmer_graph_distributed) or, fused, repeat expansion,
     renumbering and the GFA text on its device (compress_distributed)
  5. compress_distributed_split: the P lines — most of a many-assembly GFA — are printed by the ranks that own the sequences, from
     the final "<number><sign>" tokens rank 0 scatters (4 B per occurrence); the file is rank 0's text + the ranks' lines in rank order.
Ordering: when the handle runs on torch's current stream (`KmerGraph(stream=torch.cuda.current_stream().cuda_stream)`; bench.py does
this), NCCL's own ordering against that stream is all that is needed; a handle with a private stream (stream=None) is fenced by
torch.cuda.synchronize around every collective.  PyTorch is plumbing here (device buffers + NCCL); the records are produced and
consumed by the library's kernels.  The same stages with no collective library at all: ac_config.n_devices (one process, peers read
over NVLink).
"""
import ctypes as C

ENTRY_BYTES = 16
RUN_BYTES = 16


def shard_bounds(n_items, rank, world):
    """Contiguous blocks, sorted file order == rank order."""
    return n_items * rank // world, n_items * (rank + 1) // world


def _buffer(kmer_graph, name, numel, dtype, device):
    """A device buffer that lives as long as the handle and only ever grows: the exchange of a repeated build goes through the same
    addresses every time (no allocator traffic between the collectives, nothing for NCCL to set up again), and what the library
    reads asynchronously stays alive without further bookkeeping.  -> a view of exactly `numel` elements."""
    import torch
    cache = kmer_graph.__dict__.setdefault("_exchange_cache", {})
    t = cache.get(name)
    if t is None or t.numel() < numel or t.dtype != dtype or t.device != torch.device(device):
        t = torch.empty(max(1, numel + numel // 8), dtype=dtype, device=device)
        cache[name] = t
    return t[:numel]


def _exchange(kmer_graph, seq_lo, seq_hi, device, group, stats):
    """Steps 1-4 up to the imported occurrences on rank 0.  -> True on rank 0."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_cuda = torch.device(device).type == "cuda"
    # The collectives are ordered after the kernels on torch's current stream.  When the library runs on that same stream (bench.py
    # and the tests hand it over) stream order is all that is needed; a handle with a private stream needs the device to settle.
    same_stream = on_cuda and kmer_graph._h.runs_on(torch.cuda.current_stream(device).cuda_stream)

    def settle():
        if on_cuda and not same_stream:
            torch.cuda.synchronize(device)

    def clock():
        if stats is None or not on_cuda:
            return None
        e = torch.cuda.Event(enable_timing=True); e.record(); return e

    h = kmer_graph._h
    lib = h.lib
    t0 = clock()
    h.check(lib.ac_build_local(h.ptr, seq_lo, seq_hi, 1))

    # ---- k-mer buckets: sizes, then the records ----
    n = C.c_uint64()
    h.check(lib.ac_entries_count(h.ptr, C.byref(n)))
    t1 = clock()
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, torch.tensor([n.value], dtype=torch.int64, device=device), group=group)
    sizes = [int(s) for s in sizes.tolist()]
    max_n = max(sizes)
    send = _buffer(kmer_graph, "send", max(1, max_n) * ENTRY_BYTES, torch.uint8, device)
    h.check(lib.ac_entries_export(h.ptr, send.data_ptr(), max_n))
    recv = _buffer(kmer_graph, "recv", world * send.numel(), torch.uint8, device)
    dist.all_gather_into_tensor(recv, send, group=group)
    settle()
    t2 = clock()
    for r in range(world):
        if r != rank and sizes[r]:
            h.check(lib.ac_entries_merge(h.ptr, recv.data_ptr() + r * send.numel(), sizes[r]))
    t3 = clock()

    # ---- adjacency (replicated) + the owned sequences' unitig occurrences ----
    n_runs = C.c_uint64()
    h.check(lib.ac_runs_local(h.ptr, C.byref(n_runs)))
    t4 = clock()
    rsizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(rsizes, torch.tensor([n_runs.value], dtype=torch.int64, device=device), group=group)
    rsizes = [int(s) for s in rsizes.tolist()]
    max_r = max(1, max(rsizes))
    rsend = _buffer(kmer_graph, "rsend", max_r * RUN_BYTES, torch.uint8, device)
    h.check(lib.ac_runs_export(h.ptr, rsend.data_ptr(), max_r))
    rrecv = _buffer(kmer_graph, "rrecv", world * rsend.numel(), torch.uint8, device) if rank == 0 else None
    dist.gather(rsend, list(rrecv.view(world, -1).unbind(0)) if rank == 0 else None, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    settle()
    t5 = clock()
    kmer_graph._run_counts = rsizes
    if stats is not None:
        stats.update(entries_sent=n.value, entries_received=sum(sizes) - n.value, exchange_bytes_sent=n.value * ENTRY_BYTES,
                     exchange_bytes_received=(sum(sizes) - n.value) * ENTRY_BYTES, runs_total=sum(rsizes), runs_gathered_bytes=sum(rsizes) * RUN_BYTES)
        stats["_events"] = (t0, t1, t2, t3, t4, t5)
    if rank != 0:
        return False
    counts = (C.c_uint64 * world)(*rsizes)
    h.check(lib.ac_runs_import_padded(h.ptr, rrecv.data_ptr(), max_r, counts, world))
    stats is not None and stats.__setitem__("_keep", (recv, rrecv))      # the library reads these buffers asynchronously: keep them alive until the caller has synchronised
    kmer_graph._exchange_buffers = (recv, rrecv, send, rsend)
    return True


def finish_stats(stats):
    """Milliseconds between the recorded events (call after the stream has been synchronised)."""
    ev = stats.pop("_events", None)
    stats.pop("_keep", None)
    if ev and all(e is not None for e in ev):
        names = ["local_table", "exchange", "merge", "adjacency_runs", "gather_runs"]
        for i, name in enumerate(names):
            stats[name + "_ms"] = ev[i].elapsed_time(ev[i + 1])
    ev2 = stats.pop("_events2", None)
    if ev2 and all(e is not None for e in ev2):
        stats["finish_and_tokens_ms"] = ev2[0].elapsed_time(ev2[1])     # rank 0: the union graph, H/S/L text, token scatter; other ranks: waiting for it
        stats["path_lines_ms"] = ev2[1].elapsed_time(ev2[2])
    return stats


def from_kmer_graph_distributed(kmer_graph, seq_lo, seq_hi, device, group=None, stats=None):
    """-> UnitigGraph on rank 0 (the graph after from_kmer_graph), None on the other ranks.  `kmer_graph` must hold (and have uploaded) every sequence."""
    from .api import UnitigGraph
    if not _exchange(kmer_graph, seq_lo, seq_hi, device, group, stats):
        return None
    h = kmer_graph._h
    h.check(h.lib.ac_build_finish(h.ptr))
    return UnitigGraph(kmer_graph)


TOKEN_BYTES = 4


class _DeviceBytes:
    """A range of the library's device memory as something torch.as_tensor() wraps without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def upload_sharded(kmer_graph, bounds, device, group=None):
    """kmer_graph.upload() for the N-rank build: every rank copies only its own block of strands over its PCIe link (bounds[r] .. bounds[r+1]
    are rank r's sequences) and one broadcast per block brings the others' over NVLink into place.  -> bytes this rank uploaded."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    h = kmer_graph._h
    lib = h.lib
    h.check(lib.ac_upload_shard(h.ptr, bounds[rank], bounds[rank + 1]))
    on_cuda = torch.device(device).type == "cuda"
    if on_cuda and not h.runs_on(torch.cuda.current_stream(device).cuda_stream):
        torch.cuda.synchronize(device)
    work, mine = [], 0
    for r in range(world):
        ptr, n = C.c_void_p(), C.c_uint64()
        h.check(lib.ac_strand_block(h.ptr, bounds[r], bounds[r + 1], C.byref(ptr), C.byref(n)))
        if r == rank:
            mine = n.value
        if n.value == 0:
            continue
        if on_cuda:
            block = torch.as_tensor(_DeviceBytes(ptr.value, n.value), device=device)
        else:
            block = torch.frombuffer((C.c_uint8 * n.value).from_address(ptr.value), dtype=torch.uint8)
        work.append(dist.broadcast(block, src=dist.get_global_rank(group, r) if group is not None else r, group=group, async_op=True))
    for w in work:
        w.wait()
    if on_cuda and not h.runs_on(torch.cuda.current_stream(device).cuda_stream):
        torch.cuda.synchronize(device)
    return mine


class PathLines:
    """The P lines of this rank's sequences (compress_distributed(..., split_paths=True)): a view of the library's pinned buffer."""

    def __init__(self, handle):
        self._h = handle

    def view(self):
        n = C.c_uint64(); ptr = C.c_void_p()
        self._h.check(self._h.lib.ac_path_lines_data(self._h.ptr, C.byref(ptr), C.byref(n)))
        return memoryview((C.c_char * n.value).from_address(ptr.value)) if n.value else memoryview(b"")


def compress_distributed_split(kmer_graph, seq_lo, seq_hi, device, group=None, stats=None):
    """UnitigGraph.compress over the ranks with the P lines printed where the sequences live: rank 0 finishes the graph and prints H, S and
    L lines, hands every occurrence's final "<number><sign>" to the rank that owns the sequence (one scatter, 4 bytes per occurrence), and
    every rank prints the P lines of its own sequences and copies them out over its own PCIe link.
    -> (graph on rank 0 / None elsewhere, PathLines); input_assemblies.gfa = rank 0's gfa_view() + every rank's lines in rank order."""
    import torch
    import torch.distributed as dist
    from .api import UnitigGraph
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_cuda = torch.device(device).type == "cuda"
    is_root = _exchange(kmer_graph, seq_lo, seq_hi, device, group, stats)
    h = kmer_graph._h
    lib = h.lib
    rsizes = kmer_graph._run_counts
    max_r = max(1, max(rsizes))
    e0 = torch.cuda.Event(enable_timing=True) if (stats is not None and on_cuda) else None
    if e0 is not None:
        e0.record()
    graph = None
    chunk = _buffer(kmer_graph, "chunk", max_r, torch.int32, device)
    if is_root:
        h.check(lib.ac_compress_finish_split(h.ptr))
        graph = UnitigGraph(kmer_graph)
        tokens = _buffer(kmer_graph, "tokens", world * max_r, torch.int32, device)
        counts = (C.c_uint64 * world)(*rsizes)
        h.check(lib.ac_path_tokens_export(h.ptr, tokens.data_ptr(), max_r, counts, world))
        if on_cuda and not kmer_graph._h.runs_on(torch.cuda.current_stream(device).cuda_stream):
            torch.cuda.synchronize(device)
        dist.scatter(chunk, list(tokens.view(world, max_r).unbind(0)), src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    else:
        dist.scatter(chunk, None, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if on_cuda and not kmer_graph._h.runs_on(torch.cuda.current_stream(device).cuda_stream):
        torch.cuda.synchronize(device)
    e1 = torch.cuda.Event(enable_timing=True) if e0 is not None else None
    if e1 is not None:
        e1.record()
    h.check(lib.ac_path_lines_render(h.ptr, chunk.data_ptr(), rsizes[rank]))
    if stats is not None:
        e2 = torch.cuda.Event(enable_timing=True) if e0 is not None else None
        if e2 is not None:
            e2.record()
        stats["_events2"] = (e0, e1, e2)
        stats["path_tokens_bytes"] = sum(rsizes) * TOKEN_BYTES
    kmer_graph._token_buffers = (chunk,)
    return graph, PathLines(h)


def compress_distributed(kmer_graph, seq_lo, seq_hi, device, group=None, stats=None):
    """UnitigGraph.compress over the ranks: -> on rank 0 the simplified, renumbered graph with its GFA text ready, None elsewhere."""
    from .api import UnitigGraph
    if not _exchange(kmer_graph, seq_lo, seq_hi, device, group, stats):
        return None
    h = kmer_graph._h
    h.check(h.lib.ac_compress_finish(h.ptr))
    return UnitigGraph(kmer_graph)
