"""Destination-row partitioning of a graph across ranks with a halo exchange of source rows
(SURVEY §8e).  The reference has no distributed code at all (single process, single device,
``/root/reference/main.py:43``); this is the B200-box scaling path north_star asks for: one
process per GPU, nodes split into contiguous destination-row ranges balanced by work, every
rank owns the in-edges of its rows, and before each layer's aggregation the distinct remote
source rows ("halo") are fetched over NVLink.

The index logic here is device-agnostic torch (tested on CPU with the gloo backend, world_size 2 and 3,
with test-side row movers and the oracle aggregation); the row movers and the aggregation of the product
are the CUDA kernels, on the rank's local CSR whose column space is ``[local rows ‖ halo rows]``.

Forward:  X = [x_local ‖ all_to_all(x_local[send_ids])];  H = aggregate(X, local CSR)
Backward: gX = aggregate^T(gH);  gx_local = gX[:n_local] + scatter(all_to_all^T(gX[n_local:]))
          (added peer by peer in rank order, indices unique within a peer -> deterministic)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def partition_ranges(in_deg: torch.Tensor, world: int, node_cost: float = 1.0) -> List[Tuple[int, int]]:
    """Contiguous node ranges with about equal  (#in-edges + node_cost * #nodes)  each —
    the same cost the aggregation kernel's group plan balances."""
    n = int(in_deg.numel())
    cost = in_deg.to(torch.float64) + node_cost
    csum = torch.cumsum(cost, 0)
    total = float(csum[-1]) if n else 0.0
    targets = torch.tensor([total * (p + 1) / world for p in range(world - 1)], dtype=torch.float64,
                           device=in_deg.device)
    cuts = torch.searchsorted(csum, targets, right=False).tolist() if n else [0] * (world - 1)
    bounds = [0] + [min(int(c) + 1, n) for c in cuts] + [n]
    for i in range(1, len(bounds)):           # monotone, even for degenerate inputs
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[p], bounds[p + 1]) for p in range(world)]


def cyclic_relabel(edge_index: torch.Tensor, num_nodes: int, world: int):
    """Cyclic node ownership (node i -> rank i mod P) expressed as a relabelling, so that the
    rest of the pipeline keeps working on contiguous ranges:  new_id = start[i mod P] + i // P.

    Why: ranges cut on the ORIGINAL ids of a power-law graph give the hub-heavy ranks a handful
    of nodes and the tail ranks millions, and the tail ranks' rows are what every other rank
    pulls — on 8×B200 one rank then has to serve 36 % of all halo traffic (measured: the
    exchange took 5.6 ms instead of 2).  Cyclic ownership spreads hubs and tail nodes evenly, so
    compute, ingress and egress are all balanced; for R-MAT the local-edge fraction is the same
    as for range cuts because the id bits are i.i.d. across levels.
    Returns (relabelled edge_index, ranges, perm) with perm[new_id] = old_id."""
    dev = edge_index.device
    counts = [(num_nodes - p + world - 1) // world for p in range(world)]
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    st = torch.tensor(starts[:-1], dtype=torch.int64, device=dev)
    new = st[edge_index % world] + torch.div(edge_index, world, rounding_mode="floor")
    old = torch.arange(num_nodes, dtype=torch.int64, device=dev)
    perm = torch.empty(num_nodes, dtype=torch.int64, device=dev)
    perm[st[old % world] + torch.div(old, world, rounding_mode="floor")] = old
    ranges = [(starts[p], starts[p + 1]) for p in range(world)]
    return new, ranges, perm


def random_relabel(edge_index: torch.Tensor, num_nodes: int, world: int, seed: int = 1234):
    """Uniformly random node ownership expressed as a relabelling + equal contiguous ranges.

    R-MAT skews EVERY id bit, so both range cuts and cyclic (i mod P) ownership leave one rank with
    several times the edges or the egress of another (measured / simulated: 44 % of the edges on
    rank 0 for i mod 8; 36 % of the halo egress on one rank for cost-balanced ranges).  A seeded
    random permutation balances compute, ingress and egress to within a few percent, and costs no
    locality on such graphs (local-edge fraction 1/P either way).  All ranks derive the same
    permutation from the seed (CPU generator).  Returns (edge_index_new, ranges, perm) with
    perm[new_id] = old_id."""
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(num_nodes, generator=g).to(edge_index.device)     # new -> old
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(num_nodes, dtype=torch.int64, device=edge_index.device)
    new = inv[edge_index]
    base, rem = divmod(num_nodes, world)
    starts = [0]
    for p in range(world):
        starts.append(starts[-1] + base + (1 if p < rem else 0))
    return new, [(starts[p], starts[p + 1]) for p in range(world)], perm


@dataclass
class LocalPart:
    rank: int
    world: int
    ranges: List[Tuple[int, int]]
    n_local: int
    n_halo: int
    halo_gid: torch.Tensor           # int64 [n_halo] global ids of the remote source rows, ascending
    recv_splits: List[int]           # rows received from each peer (sums to n_halo)
    send_ids: torch.Tensor           # int64 [n_send] LOCAL ids of the rows each peer needs, peer-major
    send_splits: List[int]
    edge_index: torch.Tensor         # int64 [2, E_local]: src in [0, n_local+n_halo), dst in [0, n_local)
    edge_type: Optional[torch.Tensor]
    # set by restage_part(): halo rows ordered (owner, stage of first use, id) instead of (owner, id)
    stage_fractions: Optional[List[float]] = None
    recv_stage_counts: Optional[torch.Tensor] = None     # int64 [world, K] on the host: rows per (owner, stage)

    @property
    def lo(self) -> int:
        return self.ranges[self.rank][0]

    @property
    def n_src(self) -> int:
        return self.n_local + self.n_halo


def select_local(edge_index: torch.Tensor, edge_type: Optional[torch.Tensor], ranges, rank: int):
    """Communication-free half of the partition build: this rank's in-edges with sources
    renumbered into [local ‖ halo].  Returns (edge_index_local, edge_type_local, halo_gid, recv_splits)."""
    lo, hi = ranges[rank]
    dev = edge_index.device
    src, dst = edge_index[0], edge_index[1]
    mine = (dst >= lo) & (dst < hi)
    src_m, dst_m = src[mine], dst[mine] - lo
    et_m = edge_type[mine] if edge_type is not None else None
    n_local = hi - lo
    is_local = (src_m >= lo) & (src_m < hi)
    halo_gid = torch.unique(src_m[~is_local])                     # sorted ascending => grouped by owner
    src_new = torch.empty_like(src_m)
    src_new[is_local] = src_m[is_local] - lo
    src_new[~is_local] = n_local + torch.searchsorted(halo_gid, src_m[~is_local])
    starts = torch.tensor([r[0] for r in ranges] + [ranges[-1][1]], dtype=torch.int64, device=dev)
    owner_bounds = torch.searchsorted(halo_gid, starts)           # halo rows owned by peer q: [b[q], b[q+1])
    recv_splits = (owner_bounds[1:] - owner_bounds[:-1]).tolist()
    return torch.stack([src_new, dst_m]), et_m, halo_gid, recv_splits


def build_local_part(edge_index: torch.Tensor, edge_type: Optional[torch.Tensor], ranges, rank: int,
                     group=None) -> LocalPart:
    """Select this rank's in-edges, renumber sources into [local ‖ halo] and agree the send
    lists with the peers (one all_to_all of counts, one of ids)."""
    world = len(ranges)
    lo, hi = ranges[rank]
    dev = edge_index.device
    ei_local, et_m, halo_gid, recv_splits = select_local(edge_index, edge_type, ranges, rank)
    n_local = hi - lo
    n_halo = int(halo_gid.numel())
    # tell every owner which of its rows we need
    if world > 1:
        recv_cnt = torch.tensor(recv_splits, dtype=torch.int64, device=dev)
        send_cnt = torch.empty_like(recv_cnt)
        dist.all_to_all_single(send_cnt, recv_cnt, group=group)
        send_splits = send_cnt.tolist()
        send_gid = torch.empty(int(sum(send_splits)), dtype=torch.int64, device=dev)
        dist.all_to_all_single(send_gid, halo_gid, output_split_sizes=send_splits, input_split_sizes=recv_splits,
                               group=group)
        send_ids = send_gid - lo
    else:
        send_splits = [0]
        send_ids = torch.empty(0, dtype=torch.int64, device=dev)
    return LocalPart(rank=rank, world=world, ranges=list(ranges), n_local=n_local, n_halo=n_halo, halo_gid=halo_gid,
                     recv_splits=recv_splits, send_ids=send_ids, send_splits=send_splits,
                     edge_index=ei_local, edge_type=et_m)


def _pack(x_local: torch.Tensor, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rows a peer needs: the library kernel ``gmlm_gather_rows`` (raises on CPU tensors — no fallback)."""
    from .ops import gather_rows
    return gather_rows(x_local, ids, out=out)


def _unpack_add(gx: torch.Tensor, ids: torch.Tensor, rows: torch.Tensor):
    """gx[ids] += rows: the library kernel ``gmlm_scatter_add_rows`` (raises on CPU tensors — no fallback)."""
    from .ops import scatter_add_rows_
    scatter_add_rows_(gx, ids, rows)


class _HaloExchange(torch.autograd.Function):
    """x_local [n_local, F] -> X [n_local + n_halo, F] (rows of remote sources appended).

    ``pack`` / ``unpack_add`` are the row movers; the product uses the CUDA kernels above.  The CPU tests of
    the multi-rank host logic (gloo) inject torch-indexing movers of their own — test infrastructure that
    lives in tests/, like the oracle aggregation those tests use."""

    @staticmethod
    def forward(ctx, x_local, part: LocalPart, group, pack, unpack_add):
        ctx.part, ctx.group, ctx.unpack_add = part, group, unpack_add
        feat = x_local.size(1)
        X = torch.empty((part.n_src, feat), dtype=x_local.dtype, device=x_local.device)
        X[: part.n_local] = x_local
        if part.world > 1:
            send = pack(x_local, part.send_ids)
            dist.all_to_all_single(X[part.n_local:], send, output_split_sizes=part.recv_splits,
                                   input_split_sizes=part.send_splits, group=group)
        return X

    @staticmethod
    def backward(ctx, gX):
        part: LocalPart = ctx.part
        gx = gX[: part.n_local].clone()
        if part.world > 1:
            g_halo = gX[part.n_local:].contiguous()
            back = torch.empty((int(sum(part.send_splits)), gX.size(1)), dtype=gX.dtype, device=gX.device)
            dist.all_to_all_single(back, g_halo, output_split_sizes=part.send_splits,
                                   input_split_sizes=part.recv_splits, group=ctx.group)
            off = 0
            for cnt in part.send_splits:      # fixed peer order; ids unique within a peer
                if cnt:
                    ctx.unpack_add(gx, part.send_ids[off:off + cnt], back[off:off + cnt])
                off += cnt
        return gx, None, None, None, None


def halo_exchange(x_local: torch.Tensor, part: LocalPart, group=None, *, pack=None, unpack_add=None) -> torch.Tensor:
    """Autograd halo exchange over ``torch.distributed`` (NCCL on the GPUs).  Row movers default to the CUDA
    library kernels; CPU tensors raise unless a test injects its own movers."""
    return _HaloExchange.apply(x_local, part, group, pack or _pack, unpack_add or _unpack_add)



def default_stage_fractions(n_stages: int, kind: str = "fib") -> List[float]:
    """Relative block sizes of the staged forward.  ``fib``: 1, 2, 3, 5, 8, 13, ... (the first pull is the only
    exposed one, so the first block is small; the late blocks are big because most of the halo has arrived by
    then).  ``lin``: 1, 2, 3, 4, ... (a smaller last block: less compute left when the last stage lands -- for
    transports whose pull chain is about as long as the aggregation).  ``flat``: equal blocks."""
    n = max(1, n_stages)
    if kind == "lin":
        return [float(i + 1) for i in range(n)]
    if kind == "flat":
        return [1.0] * n
    fr = [1.0, 2.0]
    while len(fr) < n:
        fr.append(fr[-1] + fr[-2])
    return fr[:n]


def stage_row_cuts(rowptr: torch.Tensor, fractions) -> Tuple[List[int], torch.Tensor]:
    """Cut the CSR rows into ``len(fractions)`` consecutive blocks whose edge counts follow ``fractions``.
    Returns (row cuts [K+1], edge offset of every cut as an int64 tensor [K+1])."""
    n_rows = int(rowptr.numel()) - 1
    rowptr64 = rowptr.long()
    nnz = int(rowptr64[-1].item()) if n_rows >= 0 else 0
    tot, acc, tg = float(sum(fractions)), 0.0, []
    for f in list(fractions)[:-1]:
        acc += f
        tg.append(int(nnz * acc / tot))
    targets = torch.tensor(tg, dtype=torch.int64, device=rowptr.device)
    cuts = [0] + torch.searchsorted(rowptr64, targets, right=False).clamp(max=n_rows).tolist() + [n_rows]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    ebounds = rowptr64[torch.tensor(cuts, dtype=torch.int64, device=rowptr.device)]
    return cuts, ebounds


def first_use_stage(col: torch.Tensor, n_local: int, n_halo: int, ebounds: torch.Tensor) -> torch.Tensor:
    """For every halo row (CSR column ``n_local + h``) the first block whose edges gather it; ``K`` (= number
    of blocks) for a halo row no edge gathers."""
    K = int(ebounds.numel()) - 1
    first = torch.full((max(n_halo, 1),), K, dtype=torch.int64, device=col.device)
    if n_halo and col.numel():
        pos = torch.nonzero(col >= n_local).squeeze(1)
        blk = torch.searchsorted(ebounds, pos, right=True) - 1
        first.scatter_reduce_(0, col[pos].long() - n_local, blk, reduce="amin")
    return first[:n_halo]


def halo_first_use_stage(part: "LocalPart", live_rels: List[int], fractions) -> torch.Tensor:
    """Stage of first use of every halo row, computed from the raw local edge list exactly as
    ``build_forward_stages`` later computes it from the CSR: edges in (dst*S + slot, original order) order
    -- the order ``gmlm_csr_build`` produces -- are cut into blocks by ``stage_row_cuts`` and a halo row
    belongs to the first block that gathers it."""
    src, dst = part.edge_index[0], part.edge_index[1]
    S = max(1, len(live_rels))
    if part.edge_type is not None:
        lut = torch.full((max(live_rels) + 1,), -1, dtype=torch.int64, device=src.device)
        lut[torch.tensor(live_rels, dtype=torch.int64, device=src.device)] = torch.arange(S, device=src.device)
        key = dst * S + lut[part.edge_type]
    else:
        key = dst * S
    n_rows = part.n_local * S
    order = torch.argsort(key, stable=True)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=src.device)
    rowptr[1:] = torch.cumsum(torch.bincount(key, minlength=n_rows), 0)
    _, ebounds = stage_row_cuts(rowptr, fractions)
    return first_use_stage(src[order], part.n_local, part.n_halo, ebounds)


def restage_part(part: "LocalPart", stage: torch.Tensor, fractions, group=None) -> "LocalPart":
    """Renumber the halo rows from (owner, id) order to (owner, stage, id) order, so that the rows one stage
    needs from one owner are CONTIGUOUS both in this rank's gather matrix and in the owner's packed send
    buffer -- what a copy engine needs (contiguous peer-to-peer copies that take no SM from the aggregation).
    Owner grouping is kept, so the per-owner backward slices stay contiguous.  The send lists are agreed
    again with the owners (one all_to_all of ids), which therefore pack in the new order."""
    K = len(fractions)
    world, rank, dev = part.world, part.rank, part.edge_index.device
    n_local, n_halo = part.n_local, part.n_halo
    owner = torch.repeat_interleave(torch.arange(world, device=dev),
                                    torch.tensor(part.recv_splits, dtype=torch.int64, device=dev))
    new_order = torch.argsort(owner * (K + 1) + stage, stable=True)          # stable: ids stay ascending inside
    inv = torch.empty_like(new_order)
    inv[new_order] = torch.arange(n_halo, dtype=torch.int64, device=dev)
    halo_gid = part.halo_gid[new_order]
    src = part.edge_index[0].clone()
    is_halo = src >= n_local
    src[is_halo] = n_local + inv[src[is_halo] - n_local]
    counts = torch.zeros((world, K), dtype=torch.int64, device=dev)
    if n_halo:
        counts.view(-1).index_add_(0, owner * K + stage, torch.ones(n_halo, dtype=torch.int64, device=dev))
    if world > 1:
        send_gid = torch.empty(int(sum(part.send_splits)), dtype=torch.int64, device=dev)
        dist.all_to_all_single(send_gid, halo_gid, output_split_sizes=part.send_splits,
                               input_split_sizes=part.recv_splits, group=group)
        send_ids = send_gid - part.lo
    else:
        send_ids = part.send_ids
    return LocalPart(rank=rank, world=world, ranges=part.ranges, n_local=n_local, n_halo=n_halo, halo_gid=halo_gid,
                     recv_splits=part.recv_splits, send_ids=send_ids, send_splits=part.send_splits,
                     edge_index=torch.stack([src, part.edge_index[1]]), edge_type=part.edge_type,
                     stage_fractions=list(fractions), recv_stage_counts=counts.cpu())


def push_offset(all_splits: torch.Tensor, rank: int, owner: int) -> int:
    """Row offset, inside ``owner``'s staging area, of the gradient rows ``rank`` pushes to it.
    ``all_splits[q, p]`` = halo rows rank q gathers from rank p; the staging area is peer-major in rank
    order, i.e. laid out exactly like the owner's ``send_ids`` / ``send_splits``."""
    return int(all_splits[:rank, owner].sum())


class PeerHalo:
    """Halo exchange over NVLink peer memory (no NCCL, no pack buffers).

    The gather matrix ``X = [local ‖ halo]`` and the backward output ``gX`` live in symmetric
    memory, so every rank's kernels can address every peer's copy directly:

      forward : barrier; ONE launch of ``gmlm_gather_rows_ptr``: every halo row has its own 64-bit
                source address inside the owner's X head, so rows stream in from all peers at once
                straight over NVLink into my X tail; barrier.
      backward: barrier; ONE launch of ``gmlm_reduce_rows_ptr``: for every local row that peers
                used, the addresses of its gradient rows in those peers' gX tails (fixed peer
                order per row => deterministic, fp32 accumulate, no atomics) are summed into my
                gX head; barrier.

    Barriers are the device-side symmetric-memory barrier (≈10 µs), enqueued on the stream — no
    host synchronisation.  Measured on 2×B200: 650 GB/s per direction from inside the kernels."""

    def __init__(self, part: LocalPart, feat: int, dtype: torch.dtype, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self.part, self.feat, self.dtype = part, feat, dtype
        dev = part.edge_index.device
        world, rank = part.world, part.rank
        sizes = torch.tensor([part.n_src], dtype=torch.int64, device=dev)
        dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)
        self.max_rows = int(sizes.item())
        splits = torch.tensor(part.recv_splits, dtype=torch.int64, device=dev)
        all_splits = [torch.zeros_like(splits) for _ in range(world)]
        dist.all_gather(all_splits, splits, group=group)
        all_splits = torch.stack(all_splits).cpu()                   # [q, p] rows rank q needs from rank p
        self.X_sym = symm_mem.empty((self.max_rows, feat), dtype=dtype, device=dev)
        self.gX_sym = symm_mem.empty((self.max_rows, feat), dtype=dtype, device=dev)
        self.hx = symm_mem.rendezvous(self.X_sym, group=group.group_name)
        self.hg = symm_mem.rendezvous(self.gX_sym, group=group.group_name)
        self.X = self.X_sym[: part.n_src]
        self.gX = self.gX_sym[: part.n_src]
        self.all_splits = all_splits
        self._symm = symm_mem
        self._group = group
        self.stage_sym = None                                        # built by build_backward_push()
        esz = torch.empty(0, dtype=dtype).element_size()
        row_bytes = feat * esz
        order = sorted(range(world), key=lambda q: (q - rank) % world)      # fixed per-rank peer order
        # forward plan: one 64-bit source address per halo row (peer X head + remote local id)
        ptrs = torch.empty(part.n_halo, dtype=torch.int64, device=dev)
        key = torch.empty(part.n_halo, dtype=torch.int64, device=dev)
        off = 0
        for p in range(world):
            cnt = part.recv_splits[p]
            if cnt:
                base = self.hx.get_buffer(p, (self.max_rows, feat), dtype).data_ptr()
                ptrs[off:off + cnt] = base + (part.halo_gid[off:off + cnt] - part.ranges[p][0]) * row_bytes
                key[off:off + cnt] = torch.arange(cnt, dtype=torch.int64, device=dev) * world + (p - rank) % world
            off += cnt
        # list the rows round-robin over the owners (and start each rank on a different owner):
        # the halo list itself is grouped by owner, and walking it in order makes every rank pull
        # from the same peer at the same time — measured 5.9 ms instead of 2 on 8 GPUs
        self.fwd_order = torch.argsort(key)
        self.fwd_ptrs = ptrs[self.fwd_order].contiguous()
        # backward plan: for every local row that some peer used, the addresses of its gradient rows
        # in those peers' gX tails, grouped by row, peers in the fixed order above
        rows, addrs, keys = [], [], []
        self.bwd_seq = []
        offs = [0]
        for q in range(world):
            offs.append(offs[-1] + part.send_splits[q])
        for rank_pos, q in enumerate(order):
            cnt = part.send_splits[q]
            if not cnt or q == rank:
                continue
            n_local_q = part.ranges[q][1] - part.ranges[q][0]
            start = n_local_q + int(all_splits[q, :rank].sum())
            peer_g = self.hg.get_buffer(q, (self.max_rows, feat), dtype)
            base = peer_g.data_ptr()
            self.bwd_seq.append((q, peer_g[start:start + cnt], part.send_ids[offs[q]:offs[q] + cnt].contiguous()))
            rows.append(part.send_ids[offs[q]:offs[q] + cnt])
            addrs.append(base + (start + torch.arange(cnt, dtype=torch.int64, device=dev)) * row_bytes)
        if rows:
            rows_c, addrs_c = torch.cat(rows), torch.cat(addrs)
            srt, idx = torch.sort(rows_c, stable=True)                       # stable: keeps the peer order per row
            self.bwd_ptrs = addrs_c[idx].contiguous()
            self.bwd_rows, counts = torch.unique_consecutive(srt, return_counts=True)
            rp = torch.zeros(self.bwd_rows.numel() + 1, dtype=torch.int64, device=dev)
            rp[1:] = torch.cumsum(counts, 0)
            self.bwd_rowptr = rp.to(torch.int32)
        else:
            self.bwd_ptrs = torch.empty(0, dtype=torch.int64, device=dev)
            self.bwd_rows = torch.empty(0, dtype=torch.int64, device=dev)
            self.bwd_rowptr = torch.zeros(1, dtype=torch.int32, device=dev)

    @property
    def x_local(self) -> torch.Tensor:
        return self.X[: self.part.n_local]

    def pull_forward(self) -> torch.Tensor:
        lib = _liblib()
        self.hx.barrier()                                   # every rank's x_local is final
        if self.part.n_halo:
            tail = self.X[self.part.n_local:]
            _check(lib.gmlm_gather_rows_ptr(_p(self.fwd_ptrs), _p(self.fwd_order), _dt(self.dtype), self.feat,
                                            self.part.n_halo, _p(tail), self.feat, _st(tail.device)),
                   "gather_rows_ptr")
        self.hx.barrier()                                   # every rank is done reading
        return self.X

    # ---- pipelined backward: aggregate the halo gradients owner by owner and let each owner pull
    #      its slice over NVLink while the next slice is being computed
    def build_backward_slices(self, graph):
        """Split the transposed CSR of ``graph`` (built with keep_seg=True) into row slices: the local
        rows, then the halo rows owner by owner in this rank's rotated peer order."""
        from .graph import transpose_csr
        part, world, rank = self.part, self.part.world, self.part.rank
        if graph.seg_of_edge is None:
            raise ValueError("build_backward_slices needs RelGraph.build(..., keep_seg=True)")
        src = part.edge_index[0]
        seg = graph.seg_of_edge

        def make(a, b):
            m = (src >= a) & (src < b)
            return transpose_csr((src[m] - a).contiguous(), seg[m].contiguous(), b - a, fwd_rowptr=graph.fwd.rowptr)

        self.slice_local = make(0, part.n_local)
        offs = [0]
        for p in range(world):
            offs.append(offs[-1] + part.recv_splits[p])
        send_offs = [0]
        for q in range(world):
            send_offs.append(send_offs[-1] + part.send_splits[q])
        self.slices = []
        for k in range(1, world):
            owner = (rank + k) % world                       # the slice I compute at step k belongs to `owner`
            a, b = part.n_local + offs[owner], part.n_local + offs[owner + 1]
            src_rank = (rank - k) % world                    # ... and at step k I can pull my slice from `src_rank`
            cnt = part.send_splits[src_rank]
            pull = None
            for q, rows, ids in self.bwd_seq:
                if q == src_rank:
                    pull = (rows, ids)
            self.slices.append((make(a, b) if b > a else None, a, b, pull))
        self.comm_stream = torch.cuda.Stream(device=src.device)
        # one local staging buffer: each owner slice is fetched from the peer by a plain device-to-device
        # copy (copy engine, no SM time taken from the aggregation kernels), then added locally
        max_cnt = max([int(pl[0].size(0)) for _, _, _, pl in self.slices if pl is not None] + [1])
        self.staging = torch.empty((max_cnt, self.feat), dtype=self.dtype, device=src.device)
        return self

    def backward_pipelined(self, gh: torch.Tensor) -> torch.Tensor:
        """gX = aggregate^T(gh) slice by slice; slice k is pulled by its owner while slice k+1 is
        computed.  Summation order per local row is the fixed step order => deterministic."""
        from . import _lib
        from .ops import scatter_add_rows_, spmm
        main = torch.cuda.current_stream()
        n_local = self.part.n_local
        self.hg.barrier()                                    # peers finished reading gX of the previous step
        spmm(gh, self.slice_local, _lib.AGG_WEIGHTED, out=self.gX[:n_local])
        gx = self.gX[:n_local]
        done_local = torch.cuda.Event()
        done_local.record(main)
        self.comm_stream.wait_event(done_local)
        for csr, a, b, pull in self.slices:
            if csr is not None:
                spmm(gh, csr, _lib.AGG_WEIGHTED, out=self.gX[a:b])
            self.hg.barrier()                                # every rank has finished this step's slice
            ev = torch.cuda.Event()
            ev.record(main)
            if pull is not None:
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    stg = self.staging[: pull[0].size(0)]
                    stg.copy_(pull[0], non_blocking=True)    # contiguous remote rows over NVLink (copy engine)
                    scatter_add_rows_(gx, pull[1], stg)      # ... added into my gradient rows
        main.wait_stream(self.comm_stream)
        return gx


    # ---- pushed backward: every owner slice of the transposed aggregation is written by the
    #      aggregation kernel itself straight into the OWNER's staging area over NVLink (remote
    #      stores are posted, so the transfer rides under the kernel's gathers); after one barrier
    #      each owner adds the staged rows into its gradient locally at HBM speed.
    def build_backward_push(self, graph):
        """Needs ``build_backward_slices(graph)`` semantics (calls it when missing).  Allocates the
        symmetric staging area ``[rows peers use from me, feat]`` (peer-major, rank order) and the
        local reduce plan (per local row: its staged copies in this rank's fixed peer order)."""
        if not hasattr(self, "slices"):
            self.build_backward_slices(graph)
        part, world, rank = self.part, self.part.world, self.part.rank
        dev = part.edge_index.device
        feat, dtype = self.feat, self.dtype
        n_send = int(sum(part.send_splits))
        m = torch.tensor([max(n_send, 1)], dtype=torch.int64, device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self._group)
        self.max_stage = int(m.item())
        self.stage_sym = self._symm.empty((self.max_stage, feat), dtype=dtype, device=dev)
        self.hs = self._symm.rendezvous(self.stage_sym, group=self._group.group_name)
        offs = [0]
        for q in range(world):
            offs.append(offs[-1] + part.send_splits[q])
        # where my slice for `owner` lands in the owner's staging area: after the rows of the ranks before me
        self.push_out = []
        for k in range(1, world):
            owner = (rank + k) % world
            cnt = part.recv_splits[owner]
            off = push_offset(self.all_splits, rank, owner)
            buf = self.hs.get_buffer(owner, (self.max_stage, feat), dtype)
            self.push_out.append(buf[off:off + cnt] if cnt else None)
        # where the slice I FETCH at step k (computed for me by rank - k) lands in my own staging area
        self.fetch_dst = []
        for k in range(1, world):
            src_rank = (rank - k) % world
            cnt = part.send_splits[src_rank]
            self.fetch_dst.append(self.stage_sym[offs[src_rank]:offs[src_rank] + cnt] if cnt else None)
        # local reduce plan over my own staging area, peers in the fixed rotated order
        esz = torch.empty(0, dtype=dtype).element_size()
        row_bytes = feat * esz
        base = self.stage_sym.data_ptr()
        order = sorted(range(world), key=lambda q: (q - rank) % world)
        rows, addrs = [], []
        for q in order:
            cnt = part.send_splits[q]
            if not cnt or q == rank:
                continue
            rows.append(part.send_ids[offs[q]:offs[q] + cnt])
            addrs.append(base + (offs[q] + torch.arange(cnt, dtype=torch.int64, device=dev)) * row_bytes)
        if rows:
            rows_c, addrs_c = torch.cat(rows), torch.cat(addrs)
            srt, idx = torch.sort(rows_c, stable=True)
            self.push_ptrs = addrs_c[idx].contiguous()
            self.push_rows, counts = torch.unique_consecutive(srt, return_counts=True)
            rp = torch.zeros(self.push_rows.numel() + 1, dtype=torch.int64, device=dev)
            rp[1:] = torch.cumsum(counts, 0)
            self.push_rowptr = rp.to(torch.int32)
        else:
            self.push_ptrs = torch.empty(0, dtype=torch.int64, device=dev)
            self.push_rows = torch.empty(0, dtype=torch.int64, device=dev)
            self.push_rowptr = torch.zeros(1, dtype=torch.int32, device=dev)
        return self

    def backward_pushed(self, gh: torch.Tensor) -> torch.Tensor:
        """gX = aggregate^T(gh): local rows into my gradient, every owner's halo slice straight into that
        owner's staging area; then the staged rows are added locally (fixed peer order per row =>
        deterministic, fp32 accumulation, one rounding)."""
        from . import _lib
        from .ops import spmm
        lib = _liblib()
        n_local = self.part.n_local
        gx = self.gX[:n_local]
        main = torch.cuda.current_stream()
        self.hs.barrier()                                    # every owner has consumed last step's staged rows
        # the owner slices are bound by the NVLink store rate, not by the SMs (measured on 8 GPUs: 450 GB/s of
        # remote stores vs 0.27 ms of compute per slice), so the local rows are aggregated next to them on a
        # second stream instead of in front of them
        start = torch.cuda.Event()
        start.record(main)
        self.comm_stream.wait_event(start)
        with torch.cuda.stream(self.comm_stream):
            spmm(gh, self.slice_local, _lib.AGG_WEIGHTED, out=gx)
        for (csr, a, b, _), out in zip(self.slices, self.push_out):
            if csr is not None and out is not None:
                spmm(gh, csr, _lib.AGG_WEIGHTED, out=out)    # remote stores over NVLink
        main.wait_stream(self.comm_stream)
        self.hs.barrier()                                    # every rank's slices have landed
        n_rows = int(self.push_rows.numel())
        if n_rows:
            _check(lib.gmlm_reduce_rows_ptr(_p(gx), _dt(self.dtype), self.feat, self.feat, _p(self.push_rows),
                                            _p(self.push_rowptr), _p(self.push_ptrs), n_rows, _st(gx.device)),
                   "reduce_rows_ptr")
        return gx

    def backward_fetched(self, gh: torch.Tensor, local_last: bool = True, signals: bool = True) -> torch.Tensor:
        """Variant of ``backward_pushed`` whose halo gradients travel by copy engine instead of by the
        aggregation kernel's remote stores (measured ~450 GB/s, which bounds the pushed slices): every owner
        slice is aggregated into my own gX tail at HBM speed; its owner then fetches it with ONE contiguous
        device-to-device copy into its staging area while the next slice is being aggregated; one local reduce
        at the end (same plan, same fixed order per row as the pushed variant => bit-equal to it).

        ``local_last``: aggregate the local rows AFTER the owner slices on the main stream, so that the last
        owner slice's transfer is hidden under them instead of being exposed in front of the reduce.
        ``signals``: pairwise symmetric-memory signals (producer -> owner, on the copy stream) instead of an
        all-rank barrier after every slice: the main stream never waits for another rank inside the loop."""
        from . import _lib
        from .ops import spmm
        lib = _liblib()
        main = torch.cuda.current_stream()
        part = self.part
        world, rank = part.world, part.rank
        n_local = part.n_local
        gx = self.gX[:n_local]
        if not hasattr(self, "copy_stream"):
            self.copy_stream = torch.cuda.Stream(device=gx.device)
        self.hg.barrier()                                    # peers have fetched last step's slices from my gX tail
        if not local_last:
            start = torch.cuda.Event()
            start.record(main)
            self.comm_stream.wait_event(start)
            with torch.cuda.stream(self.comm_stream):
                spmm(gh, self.slice_local, _lib.AGG_WEIGHTED, out=gx)
        for k, ((csr, a, b, pull), dst) in enumerate(zip(self.slices, self.fetch_dst), start=1):
            owner, src_rank = (rank + k) % world, (rank - k) % world
            if csr is not None:
                spmm(gh, csr, _lib.AGG_WEIGHTED, out=self.gX[a:b])
            if signals:
                self.hg.put_signal(owner, 1)                 # my slice for `owner` is complete (release)
                with torch.cuda.stream(self.copy_stream):
                    self.hg.wait_signal(src_rank, 1)         # the slice `src_rank` computed for me is complete
                    if pull is not None and dst is not None:
                        dst.copy_(pull[0], non_blocking=True)
                continue
            self.hg.barrier()                                # every rank has finished this step's slice
            ev = torch.cuda.Event()
            ev.record(main)
            if pull is not None and dst is not None:
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(ev)
                    dst.copy_(pull[0], non_blocking=True)    # contiguous remote rows over NVLink (copy engine)
        if local_last:
            spmm(gh, self.slice_local, _lib.AGG_WEIGHTED, out=gx)
        else:
            main.wait_stream(self.comm_stream)
        main.wait_stream(self.copy_stream)
        n_rows = int(self.push_rows.numel())
        if n_rows:
            _check(lib.gmlm_reduce_rows_ptr(_p(gx), _dt(self.dtype), self.feat, self.feat, _p(self.push_rows),
                                            _p(self.push_rowptr), _p(self.push_ptrs), n_rows, _st(gx.device)),
                   "reduce_rows_ptr")
        return gx

    # ---- staged forward: the destination rows are cut into blocks of equal edge count; a halo row is
    #      pulled in the stage of the FIRST block that gathers it, on a second (high-priority) stream, so
    #      only stage 0 is exposed and every later pull runs under the previous block's aggregation.
    def build_forward_stages(self, graph, n_stages: int = 6, pull_ctas_overlapped: int = 0, fractions=None):
        """``fractions``: relative edge counts of the blocks; default grows like Fibonacci (1,2,3,5,8,13): the
        first pull is the only exposed one, so the first block is small.

        MEASURED (round 1, 10M/200M graph): not a win with SM-driven pulls, so bench.py leaves it off.  A pull
        kernel that shares the GPU with the aggregation fills its SMs' outstanding-load capacity with 3 us
        NVLink requests: sprinkled over all SMs (148 x 256 threads) it halved the aggregation's throughput
        (8 GPUs: 8.52 ms vs 8.55 ms one-shot); confined to 32 SMs (32 x 1024 threads) it only reaches
        260 GB/s (4 GPUs: 13.0 ms vs 12.0 ms).  The overlap needs the copy engines (owner-side pack +
        contiguous CE copies), see DESIGN.md §6."""
        from .graph import CSR
        part = self.part
        fwd = graph.fwd
        dev = fwd.rowptr.device
        K = max(1, int(n_stages))
        if fractions is None:
            fractions = default_stage_fractions(K)
        assert len(fractions) == K and all(f > 0 for f in fractions)
        cuts, ebounds = stage_row_cuts(fwd.rowptr, fractions)
        first = first_use_stage(fwd.col, part.n_local, part.n_halo, ebounds)   # first block that gathers each halo row
        self.fwd_stages = []
        sel_stage = first[self.fwd_order] if part.n_halo else first
        for k in range(K):
            r0, r1 = cuts[k], cuts[k + 1]
            e0, e1 = int(ebounds[k].item()), int(ebounds[k + 1].item())
            csr = None
            if r1 > r0:
                csr = CSR(rowptr=(fwd.rowptr[r0:r1 + 1] - e0).contiguous(), col=fwd.col[e0:e1], num_rows=r1 - r0,
                          hub_thresh=fwd.hub_thresh)
                csr.plan_hubs()
                csr.plan_groups()
            if part.n_halo:
                m = sel_stage == k
                ptrs, outs = self.fwd_ptrs[m].contiguous(), self.fwd_order[m].contiguous()
            else:
                ptrs = outs = torch.empty(0, dtype=torch.int64, device=dev)
            self.fwd_stages.append((csr, r0, r1, ptrs, outs))
        # halo rows no block gathers cannot exist (the halo IS the set of gathered remote rows)
        assert not part.n_halo or int((first >= K).sum().item()) == 0
        lo_pri, hi_pri = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
        self.pull_stream = torch.cuda.Stream(device=dev, priority=hi_pri)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # overlapped pulls run as a FEW 1024-thread CTAs (32 x 1024 x 128 B = 4 MB of remote loads in flight):
        # sprinkled over every SM they halved the aggregation's throughput (8 GPUs, 148 x 256 threads)
        self.pull_ctas_overlapped = int(pull_ctas_overlapped) or 32
        self.fwd_stage_rows = [int(st[3].numel()) for st in self.fwd_stages]
        # transport of the overlapped stages (k >= 1): 0 = the LDG pull kernel confined to `pull_ctas_overlapped`
        # CTAs; > 0 = that many single-warp bulk-copy (TMA) CTAs, which do not share the aggregation's load queues
        self.pull_tma_ctas = int(getattr(self, "pull_tma_ctas", 0))
        self.pull_tma_stage0 = bool(getattr(self, "pull_tma_stage0", False))
        return self

    def forward_staged(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """H = mean-aggregate([x_local ‖ halo]) with the halo pulled stage by stage under the aggregation.
        Bit-identical to ``pull_forward()`` + one whole-CSR aggregation (same rows, same per-row edge order)."""
        from . import _lib
        from .ops import spmm
        lib = _liblib()
        part = self.part
        main = torch.cuda.current_stream()
        n_rows = self.fwd_stages[-1][2]
        if out is None:
            out = torch.empty((n_rows, self.feat), dtype=self.dtype, device=self.X.device)
        tail = self.X[part.n_local:]
        timing = bool(getattr(self, "stage_timing", False))
        self.hx.barrier()                                   # every rank's x_local is final
        ready = torch.cuda.Event(enable_timing=timing)
        ready.record(main)
        self.pull_stream.wait_event(ready)
        events, agg_done = [], []
        dbg = getattr(self, "debug_skip", "")               # measurement only: "pull" / "agg" leaves that half out
        with torch.cuda.stream(self.pull_stream):
            for k, (_, _, _, ptrs, outs) in enumerate(self.fwd_stages):
                cnt = int(ptrs.numel()) if dbg != "pull" else 0
                if cnt and self.pull_tma_ctas > 0 and (k > 0 or self.pull_tma_stage0):
                    # bulk-copy engine transport: stage 0 has the GPU to itself (one CTA per SM), later stages
                    # run under an aggregation kernel from a few single-warp CTAs
                    warps, kb, rb = getattr(self, "pull_tma_shape", (0, 0, 0))      # (warps, ring KiB, rows per batch)
                    if k == 0:
                        _check(lib.gmlm_gather_rows_ptr_tma(_p(ptrs), _p(outs), _dt(self.dtype), self.feat, cnt, _p(tail),
                                                            self.feat, 0, 0, 0, 0, _st(tail.device)), "gather_rows_ptr_tma")
                    else:
                        _check(lib.gmlm_gather_rows_ptr_tma(_p(ptrs), _p(outs), _dt(self.dtype), self.feat, cnt, _p(tail),
                                                            self.feat, self.pull_tma_ctas, warps, kb, rb,
                                                            _st(tail.device)), "gather_rows_ptr_tma")
                elif cnt:
                    # stage 0 has the GPU to itself; later stages share it with an aggregation kernel
                    lib.gmlm_set_tuning(b"halo_pull_ctas", 0 if k == 0 else self.pull_ctas_overlapped)
                    lib.gmlm_set_tuning(b"halo_pull_threads", 0 if k == 0 else 1024)
                    _check(lib.gmlm_gather_rows_ptr(_p(ptrs), _p(outs), _dt(self.dtype), self.feat, cnt, _p(tail),
                                                    self.feat, _st(tail.device)), "gather_rows_ptr")
                ev = torch.cuda.Event(enable_timing=timing)
                ev.record(self.pull_stream)
                events.append(ev)
            lib.gmlm_set_tuning(b"halo_pull_ctas", 0)
            lib.gmlm_set_tuning(b"halo_pull_threads", 0)
        # the blocks depend on their halo stage, not on each other (disjoint output rows): alternate them over two
        # streams so that the tail of block k (partial last wave of its rows / chunk / final kernels) is filled
        # by the start of block k+1 -- measured on 2 GPUs, 4 serial blocks cost 12 % over one whole aggregation
        if not hasattr(self, "agg_streams"):
            self.agg_streams = [torch.cuda.Stream(device=self.X.device), torch.cuda.Stream(device=self.X.device)]
        two = bool(getattr(self, "overlap_blocks", True)) and not timing and len(self.fwd_stages) > 1
        if two:
            for st in self.agg_streams:
                st.wait_event(ready)
        for k, ((csr, r0, r1, _, _), ev) in enumerate(zip(self.fwd_stages, events)):
            st = self.agg_streams[k % 2] if two else main
            st.wait_event(ev)
            if csr is not None and dbg != "agg":
                with torch.cuda.stream(st):
                    spmm(self.X, csr, _lib.AGG_MEAN, out=out[r0:r1])
            if timing:
                d = torch.cuda.Event(enable_timing=True)
                d.record(main)
                agg_done.append(d)
        if two:
            for st in self.agg_streams:
                main.wait_stream(st)
        self.hx.barrier()                                   # every rank is done reading
        if timing:
            self.last_timeline = (ready, events, agg_done)
        return out

    def timeline_ms(self):
        """(pull-stage end times, block end times) of the last ``forward_staged`` call, in ms after its first
        barrier; needs ``stage_timing = True`` and a device synchronisation by the caller."""
        t0, pulls, aggs = self.last_timeline
        return [round(t0.elapsed_time(e), 3) for e in pulls], [round(t0.elapsed_time(e), 3) for e in aggs]

    # ---- packed forward: owners pack the rows every peer needs into a contiguous symmetric send buffer
    #      (one local gather), peers fetch their (owner, stage) ranges with plain device-to-device copies
    #      -- copy engines, no SM taken from the aggregation, unlike the SM-driven staged pull above --
    #      and block k of the aggregation starts when stage k has landed.  Needs a part from restage_part().
    def build_forward_packed(self, graph):
        from .graph import CSR
        part = self.part
        if part.recv_stage_counts is None:
            raise ValueError("build_forward_packed needs a LocalPart from restage_part()")
        world, rank = part.world, part.rank
        fwd = graph.fwd
        dev = fwd.rowptr.device
        feat, dtype = self.feat, self.dtype
        K = len(part.stage_fractions)
        cuts, ebounds = stage_row_cuts(fwd.rowptr, part.stage_fractions)
        n_send = int(sum(part.send_splits))
        m = torch.tensor([max(n_send, 1)], dtype=torch.int64, device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self._group)
        self.max_send = int(m.item())
        self.send_sym = self._symm.empty((self.max_send, feat), dtype=dtype, device=dev)
        self.hp = self._symm.rendezvous(self.send_sym, group=self._group.group_name)
        self.send_buf = self.send_sym[:n_send] if n_send else None
        tail_off = [0]
        for o in range(world):
            tail_off.append(tail_off[-1] + part.recv_splits[o])
        cnt = part.recv_stage_counts                                     # [owner, stage]
        self.packed_stages = []
        for k in range(K):
            r0, r1 = cuts[k], cuts[k + 1]
            e0, e1 = int(ebounds[k].item()), int(ebounds[k + 1].item())
            csr = None
            if r1 > r0:
                csr = CSR(rowptr=(fwd.rowptr[r0:r1 + 1] - e0).contiguous(), col=fwd.col[e0:e1], num_rows=r1 - r0,
                          hub_thresh=fwd.hub_thresh)
                csr.plan_hubs()
                csr.plan_groups()
            copies = []
            for j in range(1, world):                                    # rotated owner order: no two ranks start on the same peer
                o = (rank + j) % world
                c = int(cnt[o, k])
                if not c:
                    continue
                within = int(cnt[o, :k].sum())
                dst = self.X[part.n_local + tail_off[o] + within: part.n_local + tail_off[o] + within + c]
                base = push_offset(self.all_splits, rank, o)            # my rows inside owner o's send buffer
                src = self.hp.get_buffer(o, (self.max_send, feat), dtype)[base + within: base + within + c]
                copies.append((dst, src))
            self.packed_stages.append((csr, r0, r1, copies))
        if not hasattr(self, "copy_stream"):
            self.copy_stream = torch.cuda.Stream(device=dev)
        return self

    def forward_packed(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """H = mean-aggregate([x_local ‖ halo]); bit-identical to the one-shot forward."""
        from . import _lib
        from .ops import gather_rows, spmm
        part = self.part
        main = torch.cuda.current_stream()
        n_rows = self.packed_stages[-1][2]
        if out is None:
            out = torch.empty((n_rows, self.feat), dtype=self.dtype, device=self.X.device)
        if self.send_buf is not None:
            gather_rows(self.x_local, part.send_ids, out=self.send_buf)     # pack: [peer][stage][id]
        self.hp.barrier()                                                   # every owner's send buffer is final
        ready = torch.cuda.Event()
        ready.record(main)
        self.copy_stream.wait_event(ready)
        events = []
        with torch.cuda.stream(self.copy_stream):
            for _, _, _, copies in self.packed_stages:
                for dst, src in copies:
                    dst.copy_(src, non_blocking=True)                       # contiguous rows over NVLink (copy engine)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                events.append(ev)
        if not hasattr(self, "agg_streams"):
            self.agg_streams = [torch.cuda.Stream(device=self.X.device), torch.cuda.Stream(device=self.X.device)]
        two = bool(getattr(self, "overlap_blocks", True)) and len(self.packed_stages) > 1
        if two:                                                             # blocks alternate over two streams: see
            for st in self.agg_streams:                                     # forward_staged
                st.wait_event(ready)
        for k, ((csr, r0, r1, _), ev) in enumerate(zip(self.packed_stages, events)):
            st = self.agg_streams[k % 2] if two else main
            st.wait_event(ev)
            if csr is not None:
                with torch.cuda.stream(st):
                    spmm(self.X, csr, _lib.AGG_MEAN, out=out[r0:r1])
        if two:
            for st in self.agg_streams:
                main.wait_stream(st)
        self.hp.barrier()                                                   # every peer has fetched its rows
        return out

    def pull_backward(self) -> torch.Tensor:
        """gX holds the transposed aggregation's output; returns grad wrt the local rows."""
        lib = _liblib()
        self.hg.barrier()                                   # every rank's gX is final
        gx = self.gX[: self.part.n_local]
        n_rows = int(self.bwd_rows.numel())
        if n_rows:
            _check(lib.gmlm_reduce_rows_ptr(_p(gx), _dt(self.dtype), self.feat, self.feat, _p(self.bwd_rows),
                                            _p(self.bwd_rowptr), _p(self.bwd_ptrs), n_rows, _st(gx.device)),
                   "reduce_rows_ptr")
        self.hg.barrier()
        return gx


def _liblib():
    from . import _lib
    return _lib.load()


def _check(rc, what):
    from . import _lib
    _lib.check(rc, what)


def _p(t):
    from .graph import _ptr
    return _ptr(t)


def _st(dev):
    from .graph import _stream
    return _stream(dev)


def _dt(dtype):
    from . import _lib
    return {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[dtype]
