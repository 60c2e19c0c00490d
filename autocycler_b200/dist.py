"""Multi-GPU form of UnitigGraph.from_kmer_graph / UnitigGraph.compress: one process per GPU under torch.distributed.

Every rank stages and uploads ALL sequences (they are small: 1 byte per base) and owns a contiguous block of them.
  1. local k-mer table over the owned sequences                       (parallel: the dominant insert work is split N ways)
  2. ONE data-path collective: all-gather of the deduplicated local tables (16 B per distinct k-mer) over NVLink,
     merged into every rank's table (counts add, first/last flags OR, the slot keeps the smallest occurrence, which
     names the k-mer identically on every rank)
  3. adjacency on the now global table (replicated), unitig occurrences along the owned sequences (parallel)
  4. gather of the occurrences (16 B each) to rank 0, which builds unitigs / links / seed order over all of them and
     continues exactly like the single-GPU path: host graph (from_kmer_graph_distributed) or, fused, repeat expansion,
     renumbering and the GFA text on its device (compress_distributed).
PyTorch is plumbing here (device buffers + NCCL); the records are produced and consumed by the library's kernels.
"""
import ctypes as C

ENTRY_BYTES = 16
RUN_BYTES = 16


def shard_bounds(n_items, rank, world):
    """Contiguous blocks, sorted file order == rank order."""
    return n_items * rank // world, n_items * (rank + 1) // world


def _exchange(kmer_graph, seq_lo, seq_hi, device, group, stats):
    """Steps 1-4 up to the imported occurrences on rank 0.  -> True on rank 0."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_cuda = torch.device(device).type == "cuda"
    # The collectives are ordered after the kernels on torch's current stream.  When the library runs on that same stream (bench.py
    # and the tests hand it over) stream order is all that is needed; a handle with a private stream needs the device to settle.
    same_stream = on_cuda and kmer_graph._h.stream == torch.cuda.current_stream(device).cuda_stream

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
    send = torch.empty(max(1, max_n) * ENTRY_BYTES, dtype=torch.uint8, device=device)
    h.check(lib.ac_entries_export(h.ptr, send.data_ptr(), max_n))
    recv = torch.empty(world * send.numel(), dtype=torch.uint8, device=device)
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
    rsend = torch.empty(max_r * RUN_BYTES, dtype=torch.uint8, device=device)
    h.check(lib.ac_runs_export(h.ptr, rsend.data_ptr(), max_r))
    rrecv = torch.empty(world * rsend.numel(), dtype=torch.uint8, device=device) if rank == 0 else None
    dist.gather(rsend, list(rrecv.view(world, -1).unbind(0)) if rank == 0 else None, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    settle()
    t5 = clock()
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
    return stats


def from_kmer_graph_distributed(kmer_graph, seq_lo, seq_hi, device, group=None, stats=None):
    """-> UnitigGraph on rank 0 (the graph after from_kmer_graph), None on the other ranks.  `kmer_graph` must hold (and have uploaded) every sequence."""
    from .api import UnitigGraph
    if not _exchange(kmer_graph, seq_lo, seq_hi, device, group, stats):
        return None
    h = kmer_graph._h
    h.check(h.lib.ac_build_finish(h.ptr))
    return UnitigGraph(kmer_graph)


def compress_distributed(kmer_graph, seq_lo, seq_hi, device, group=None, stats=None):
    """UnitigGraph.compress over the ranks: -> on rank 0 the simplified, renumbered graph with its GFA text ready, None elsewhere."""
    from .api import UnitigGraph
    if not _exchange(kmer_graph, seq_lo, seq_hi, device, group, stats):
        return None
    h = kmer_graph._h
    h.check(h.lib.ac_compress_finish(h.ptr))
    return UnitigGraph(kmer_graph)
