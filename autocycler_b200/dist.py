"""Multi-GPU form of UnitigGraph.from_kmer_graph: one process per GPU under torch.distributed.

Every rank stages and uploads ALL sequences (they are small: 1 byte per base) and owns a contiguous block of them.
  1. local k-mer table over the owned sequences                       (parallel: the dominant insert work is split N ways)
  2. ONE data-path collective: all-gather of the deduplicated local tables (16 B per distinct k-mer) over NVLink,
     merged into every rank's table (counts add, first/last flags OR, the entry keeps the smallest occurrence, which
     names the k-mer identically on every rank)
  3. adjacency on the now global table (replicated), unitig occurrences along the owned sequences (parallel)
  4. gather of the occurrences (32 B each) to rank 0, which builds unitigs / links / seed order over all of them and
     continues exactly like the single-GPU path (host graph, simplify, GFA).
PyTorch is plumbing here (device buffers + NCCL); the records are produced and consumed by the library's kernels.
"""
import ctypes as C

ENTRY_BYTES = 16
RUN_BYTES = 32


def shard_bounds(n_items, rank, world):
    """Contiguous blocks, sorted file order == rank order."""
    return n_items * rank // world, n_items * (rank + 1) // world


def from_kmer_graph_distributed(kmer_graph, seq_lo, seq_hi, device, group=None, stats=None):
    """-> UnitigGraph on rank 0, None on the other ranks.  `kmer_graph` must hold (and have uploaded) every sequence."""
    import torch
    import torch.distributed as dist
    from .api import UnitigGraph

    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def settle():
        # The library may run on its own stream: make every collective's result visible before a kernel reads it.
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)

    h = kmer_graph._h
    lib = h.lib
    h.check(lib.ac_build_local(h.ptr, seq_lo, seq_hi, 1))

    # ---- k-mer buckets: sizes, then the records ----
    n = C.c_uint64()
    h.check(lib.ac_entries_count(h.ptr, C.byref(n)))
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n.value], dtype=torch.int64, device=device), group=group)
    sizes = [int(s.item()) for s in sizes]
    max_n = max(sizes)
    send = torch.zeros(max(1, max_n) * ENTRY_BYTES, dtype=torch.uint8, device=device)
    h.check(lib.ac_entries_export(h.ptr, send.data_ptr(), max_n))
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    settle()
    for r in range(world):
        if r != rank and sizes[r]:
            h.check(lib.ac_entries_merge(h.ptr, recv[r].data_ptr(), sizes[r]))
    if stats is not None:
        stats["entries_sent"] = n.value
        stats["entries_received"] = sum(sizes) - n.value

    # ---- adjacency (replicated) + the owned sequences' unitig occurrences ----
    n_runs = C.c_uint64()
    h.check(lib.ac_runs_local(h.ptr, C.byref(n_runs)))
    rsizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(rsizes, torch.tensor([n_runs.value], dtype=torch.int64, device=device), group=group)
    rsizes = [int(s.item()) for s in rsizes]
    max_r = max(rsizes)
    rsend = torch.zeros(max(1, max_r) * RUN_BYTES, dtype=torch.uint8, device=device)
    h.check(lib.ac_runs_export(h.ptr, rsend.data_ptr(), max_r))
    rrecv = [torch.empty_like(rsend) for _ in range(world)] if rank == 0 else None
    dist.gather(rsend, rrecv, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    settle()
    if rank != 0:
        return None
    all_runs = torch.cat([rrecv[r][:rsizes[r] * RUN_BYTES] for r in range(world)])   # rank order == coordinate order
    settle()
    h.check(lib.ac_runs_import(h.ptr, all_runs.data_ptr(), sum(rsizes)))
    h.check(lib.ac_build_finish(h.ptr))
    if stats is not None:
        stats["runs_total"] = sum(rsizes)
    return UnitigGraph(kmer_graph)
