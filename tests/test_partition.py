"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): destination-row partitioning, halo
index lists and the halo exchange forward/backward.  The aggregation in these tests is the
ORACLE (the product kernel needs a GPU); what is under test is that partition + exchange +
local aggregation reproduces the single-partition result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gmlm_b200.partition import build_local_part, halo_exchange, partition_ranges
from gmlm_b200 import synth
from oracle import edge_type_bucket_ref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _aggregate_all_rel(x, ei, et, n_dst, rels):
    outs = []
    for r in rels:
        m = et == r
        src, dst = ei[0, m], ei[1, m]
        x_j = x.index_select(0, src)
        s = torch.zeros((n_dst, x.size(1)), dtype=x.dtype).index_add_(0, dst, x_j)
        c = torch.zeros(n_dst, dtype=x.dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=x.dtype))
        outs.append(s / c.clamp(min=1).unsqueeze(-1))
    return torch.cat(outs, 1)


def _cpu_pack(x_local, ids, out=None):
    """Test-side row movers for the gloo runs (the product's are CUDA kernels and refuse CPU tensors)."""
    return x_local.index_select(0, ids)


def _cpu_unpack_add(gx, ids, rows):
    gx.index_add_(0, ids, rows)


def _worker(rank, world, port, n, e, feat, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ei = synth.rmat_edges(n, e, seed=5)
        et = edge_type_bucket_ref(ei, n)
        rels = sorted(set(et.tolist()))
        x = torch.randn(n, feat, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
        gh_full = torch.randn(n, len(rels) * feat, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
        in_deg = torch.bincount(ei[1], minlength=n)
        ranges = partition_ranges(in_deg, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        part = build_local_part(ei, et, ranges, rank)
        lo, hi = ranges[rank]
        # every edge whose dst is mine is kept, with sources renumbered into [local ‖ halo]
        assert part.edge_index.size(1) == int(((ei[1] >= lo) & (ei[1] < hi)).sum())
        assert part.edge_index.size(1) == 0 or int(part.edge_index[0].max()) < part.n_src
        assert sum(part.recv_splits) == part.n_halo and part.recv_splits[rank] == 0

        x_local = x[lo:hi].clone().requires_grad_(True)
        with pytest.raises(Exception):                                   # product movers: CUDA only, no CPU fallback
            halo_exchange(x_local, part)
        X = halo_exchange(x_local, part, pack=_cpu_pack, unpack_add=_cpu_unpack_add)
        assert torch.equal(X[: part.n_local], x[lo:hi])
        assert torch.equal(X[part.n_local:], x[part.halo_gid])           # the right remote rows arrived

        h_local = _aggregate_all_rel(X, part.edge_index, part.edge_type, part.n_local, rels)
        h_full = _aggregate_all_rel(x, ei, et, n, rels)
        assert torch.allclose(h_local, h_full[lo:hi], rtol=0, atol=1e-12)

        # backward: local grads + halo grads returned to their owners == global gradient
        h_local.backward(gh_full[lo:hi])
        xg = x.clone().requires_grad_(True)
        _aggregate_all_rel(xg, ei, et, n, rels).backward(gh_full)
        assert torch.allclose(x_local.grad, xg.grad[lo:hi], rtol=0, atol=1e-12)

        # restaged halo order (owner, stage of first use, id): same rows, same result, contiguous (owner, stage)
        # ranges -- the layout the copy-engine forward (PeerHalo.forward_packed) relies on
        from gmlm_b200.partition import (default_stage_fractions, first_use_stage, halo_first_use_stage, restage_part,
                                         stage_row_cuts)
        fr = default_stage_fractions(4)
        stage = halo_first_use_stage(part, rels, fr)
        part2 = restage_part(part, stage, fr)
        assert part2.recv_splits == part.recv_splits and part2.send_splits == part.send_splits
        assert torch.equal(torch.sort(part2.halo_gid).values, part.halo_gid)
        assert int(part2.recv_stage_counts.sum()) == part.n_halo
        assert part2.recv_stage_counts.sum(1).tolist() == part.recv_splits
        x_local2 = x[lo:hi].clone().requires_grad_(True)
        X2 = halo_exchange(x_local2, part2, pack=_cpu_pack, unpack_add=_cpu_unpack_add)   # owners pack in the NEW order
        assert torch.equal(X2[part2.n_local:], x[part2.halo_gid])
        # same global source behind every local edge
        gid_of = torch.cat([torch.arange(lo, hi), part.halo_gid])
        gid_of2 = torch.cat([torch.arange(lo, hi), part2.halo_gid])
        assert torch.equal(gid_of[part.edge_index[0]], gid_of2[part2.edge_index[0]])
        assert torch.equal(part.edge_index[1], part2.edge_index[1])
        h2 = _aggregate_all_rel(X2, part2.edge_index, part2.edge_type, part2.n_local, rels)
        assert torch.equal(h2, h_local.detach())
        h2.backward(gh_full[lo:hi])
        assert torch.allclose(x_local2.grad, xg.grad[lo:hi], rtol=0, atol=1e-12)
        # layout: within every owner group the stages are ascending and contiguous, ids ascending inside a stage,
        # and the stage of a row (recomputed on the renumbered part) is the block of its first use
        stage2 = halo_first_use_stage(part2, rels, fr)
        off = 0
        for o in range(world):
            seg = stage2[off:off + part2.recv_splits[o]]
            assert torch.equal(seg, torch.sort(seg).values)
            assert torch.bincount(seg, minlength=len(fr)).tolist() == part2.recv_stage_counts[o].tolist()
            gseg = part2.halo_gid[off:off + part2.recv_splits[o]]
            for k in range(len(fr)):
                ids = gseg[seg == k]
                assert torch.equal(ids, torch.sort(ids).values)
            assert part2.recv_splits[o] == 0 or (int(gseg.min()) >= ranges[o][0] and int(gseg.max()) < ranges[o][1])
            off += part2.recv_splits[o]
        q.put((rank, "ok", part.n_halo))
    except Exception as ex:  # surface the failure in the parent
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 600, 9000, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"


def test_partition_ranges_balance_and_degenerate():
    in_deg = torch.tensor([100, 0, 0, 0, 50, 50, 0, 0])
    r = partition_ranges(in_deg, 2)
    assert r[0][0] == 0 and r[-1][1] == 8 and r[0][1] == r[1][0]
    cost = in_deg.double() + 1
    halves = [float(cost[a:b].sum()) for a, b in r]
    assert abs(halves[0] - halves[1]) <= float(cost.max())
    # more ranks than nodes: empty ranges are allowed, coverage is kept
    r = partition_ranges(torch.tensor([3, 1]), 4)
    assert r[0][0] == 0 and r[-1][1] == 2 and all(a <= b for a, b in r)
    assert sum(b - a for a, b in r) == 2


def test_cyclic_relabel_is_a_permutation_and_balances():
    from gmlm_b200.partition import cyclic_relabel
    n, world = 1003, 4
    ei = synth.rmat_edges(n, 20000, seed=3)
    new, ranges, perm = cyclic_relabel(ei, n, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert sorted(perm.tolist()) == list(range(n))
    assert torch.equal(perm[new], ei)                        # new -> old round trip
    for p, (lo, hi) in enumerate(ranges):
        assert bool((perm[lo:hi] % world == p).all())        # rank p owns the ids congruent to p
        assert abs((hi - lo) - n / world) <= 1
    # degrees are invariant under the relabelling, so edge typing is too
    assert torch.equal(edge_type_bucket_ref(new, n), edge_type_bucket_ref(ei, n))


def test_random_relabel_is_a_seeded_permutation_and_balances_edges():
    from gmlm_b200.partition import random_relabel
    n, world = 20011, 8
    ei = synth.rmat_edges(n, 400000, seed=3)
    new, ranges, perm = random_relabel(ei, n, world)
    new2, ranges2, perm2 = random_relabel(ei, n, world)
    assert torch.equal(new, new2) and ranges == ranges2 and torch.equal(perm, perm2)   # every rank derives the same
    assert sorted(perm.tolist()) == list(range(n))
    assert torch.equal(perm[new], ei)
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(abs((b - a) - n / world) <= 1 for a, b in ranges)
    assert torch.equal(edge_type_bucket_ref(new, n), edge_type_bucket_ref(ei, n))
    starts = torch.tensor([r[0] for r in ranges] + [n])
    own = torch.searchsorted(starts, new[1], right=True) - 1
    per_rank = torch.bincount(own, minlength=world).double()
    assert float(per_rank.max() / per_rank.mean()) < 1.5      # i mod P ownership gives 3.5x on this graph


# ----------------------------------------------------------------------------- staged forward / pushed backward
# Index logic of the overlapped exchange (PeerHalo.build_forward_stages / build_backward_push); the
# kernels and the symmetric-memory plumbing themselves are covered on real GPUs by tests/multi_gpu_check.py.
def _csr_by_dst(ei, n_dst):
    order = torch.argsort(ei[1], stable=True)
    col = ei[0][order].to(torch.int32)
    rowptr = torch.zeros(n_dst + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(ei[1], minlength=n_dst), 0)
    return rowptr.to(torch.int32), col


@pytest.mark.parametrize("world,stages", [(2, 1), (3, 4), (4, 6)])
def test_stage_plan_covers_every_halo_row_before_its_first_use(world, stages):
    from gmlm_b200.partition import (default_stage_fractions, first_use_stage, random_relabel, select_local,
                                     stage_row_cuts)
    n, e = 3000, 40000
    ei = synth.rmat_edges(n, e, seed=3)
    ei, ranges, _ = random_relabel(ei, n, world)
    for rank in range(world):
        ei_l, _, halo_gid, recv_splits = select_local(ei, None, ranges, rank)
        n_local, n_halo = ranges[rank][1] - ranges[rank][0], int(halo_gid.numel())
        rowptr, col = _csr_by_dst(ei_l, n_local)
        fr = default_stage_fractions(stages)
        assert len(fr) == stages and fr == sorted(fr)
        cuts, eb = stage_row_cuts(rowptr, fr)
        assert cuts[0] == 0 and cuts[-1] == n_local and cuts == sorted(cuts) and len(cuts) == stages + 1
        assert int(eb[0]) == 0 and int(eb[-1]) == col.numel()
        first = first_use_stage(col, n_local, n_halo, eb)
        assert first.numel() == n_halo and (n_halo == 0 or int(first.max()) < stages)   # every halo row is gathered
        # brute force: the block of every edge is >= the stage its halo row arrives in, with equality somewhere
        blk_of_edge = torch.searchsorted(eb, torch.arange(col.numel()), right=True) - 1
        seen = torch.full((n_halo,), stages, dtype=torch.int64)
        for pos in torch.nonzero(col >= n_local).squeeze(1).tolist():
            h = int(col[pos]) - n_local
            assert int(blk_of_edge[pos]) >= int(first[h])
            seen[h] = min(int(seen[h]), int(blk_of_edge[pos]))
        assert torch.equal(seen, first)
        # block edge counts follow the fractions to within one row's worth of edges
        max_row = int((rowptr[1:] - rowptr[:-1]).max())
        tot = float(sum(fr))
        acc = 0.0
        for k in range(stages - 1):
            acc += fr[k]
            assert abs(int(eb[k + 1]) - col.numel() * acc / tot) <= max_row + 1


@pytest.mark.parametrize("world", [2, 3, 5])
def test_push_offsets_tile_every_owners_staging_area(world):
    """Rank p's slice for owner o lands at push_offset(all_splits, p, o); over all p these ranges must tile
    o's staging area in exactly the layout of o's send lists (peer-major, rank order, ascending ids)."""
    from gmlm_b200.partition import push_offset, random_relabel, select_local
    n, e = 2000, 30000
    ei = synth.rmat_edges(n, e, seed=8)
    ei, ranges, _ = random_relabel(ei, n, world)
    halo = [select_local(ei, None, ranges, r) for r in range(world)]
    all_splits = torch.tensor([h[3] for h in halo], dtype=torch.int64)          # [q, p] rows q gathers from p
    assert int(torch.diagonal(all_splits).sum()) == 0
    for owner in range(world):
        lo = ranges[owner][0]
        send_splits = all_splits[:, owner].tolist()                             # what build_local_part agrees on
        staged = torch.full((sum(send_splits),), -1, dtype=torch.int64)
        for p in range(world):
            cnt = send_splits[p]
            off = push_offset(all_splits, p, owner)
            assert off == sum(send_splits[:p])
            gid = halo[p][2]                                                    # p's halo rows, ascending, grouped by owner
            start = int(all_splits[p, :owner].sum())
            staged[off:off + cnt] = gid[start:start + cnt] - lo                 # local id (at the owner) of each pushed row
        assert int((staged < 0).sum()) == 0
        assert int(staged.min()) >= 0 and int(staged.max()) < ranges[owner][1] - lo
        # within one sender the ids ascend, so the owner's send_ids (same construction) match position by position
        off = 0
        for p in range(world):
            seg = staged[off:off + send_splits[p]]
            assert torch.equal(seg, torch.sort(seg).values) and seg.unique().numel() == seg.numel()
            off += send_splits[p]


# ----------------------------------------------------------------------------- partitioned GraphNorm (A7 over ranks)
class _EmulatedGraphNormKernels:
    """fp64 restatement of the four kernel entry points behind A7, formula by formula as in
    gmlm_b200/csrc/graphnorm.cu (sums and the row count they cover passed in), so that the rank logic of
    gmlm_b200.dist_norm can be checked on CPU: test infrastructure, like the oracle."""

    @staticmethod
    def _gelu_grad(n):
        cdf = 0.5 * (1.0 + torch.erf(n / 2.0 ** 0.5))
        return cdf + n * torch.exp(-0.5 * n * n) / (2.0 * torch.pi) ** 0.5

    @staticmethod
    def colstats(x):
        return x.double().sum(0), (x.double() ** 2).sum(0)

    @staticmethod
    def fwd(x, colsum, colsq, weight, bias, mean_scale, eps, fuse_gelu, stat_rows):
        n = stat_rows or x.size(0)
        mu = colsum / n
        a = mean_scale.double()
        var = (colsq / n - mu * mu * (2 * a - a * a)).clamp(min=0)
        rstd = 1.0 / torch.sqrt(var + eps)
        y = (x.double() - mu * a) * (weight.double() * rstd) + bias.double()
        if fuse_gelu:
            y = torch.nn.functional.gelu(y)
        return y.to(x.dtype), mu, rstd

    @classmethod
    def _dn(cls, x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu):
        oh = (x.double() - mean * mean_scale.double()) * rstd
        dn = gy.double()
        if fuse_gelu:
            dn = dn * cls._gelu_grad(oh * weight.double() + bias.double())
        return oh, dn

    @classmethod
    def bwd_stats(cls, x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu):
        oh, dn = cls._dn(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu)
        return dn.sum(0), (dn * oh).sum(0)

    @classmethod
    def bwd_apply(cls, x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, s1, s2, need_gx, stat_rows):
        n = stat_rows or x.size(0)
        oh, dn = cls._dn(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu)
        w, a = weight.double(), mean_scale.double()
        sum_do = w * rstd * (s1 - s2 * rstd * mean * (1.0 - a))
        gx = (w * rstd) * (dn - oh * (s2 / n)) - a * sum_do / n
        return (gx.to(x.dtype) if need_gx else None), s2.clone(), s1.clone(), -mean * sum_do


def _gn_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gmlm_b200.dist_norm import partitioned_graph_norm
        from oracle import GraphNormRef
        n, c = 157, 12
        torch.manual_seed(0)                                      # replicated parameters: same draw on every rank
        g = torch.Generator().manual_seed(4)
        x = torch.randn(n, c, dtype=torch.float64, generator=g) * 2 + 3
        gout = torch.randn(n, c, dtype=torch.float64, generator=g)
        cuts = [0, 40, 41, 41][: world] + [n]                     # uneven shards: a single row, and (world 4) NO rows
        lo, hi = cuts[rank], cuts[rank + 1]
        for fuse_gelu in (False, True):
            ref = GraphNormRef(c).double()
            with torch.no_grad():
                ref.weight.uniform_(0.5, 1.5)
                ref.bias.uniform_(-0.5, 0.5)
                ref.mean_scale.uniform_(0.2, 1.2)
            x64 = x.clone().requires_grad_(True)
            y_ref = ref(x64)
            if fuse_gelu:
                y_ref = torch.nn.functional.gelu(y_ref)
            y_ref.backward(gout)
            w = ref.weight.detach().clone().requires_grad_(True)
            b = ref.bias.detach().clone().requires_grad_(True)
            ms = ref.mean_scale.detach().clone().requires_grad_(True)
            xl = x[lo:hi].clone().requires_grad_(True)
            y = partitioned_graph_norm(xl, w, b, ms, n, ref.eps, fuse_gelu, backend=_EmulatedGraphNormKernels)
            y.backward(gout[lo:hi])
            assert torch.allclose(y, y_ref[lo:hi].detach(), rtol=1e-10, atol=1e-12)
            assert torch.allclose(xl.grad, x64.grad[lo:hi], rtol=1e-9, atol=1e-11)
            # parameter gradients are the WHOLE-GRAPH gradients on every rank (no further all-reduce)
            assert torch.allclose(w.grad, ref.weight.grad, rtol=1e-9, atol=1e-11)
            assert torch.allclose(b.grad, ref.bias.grad, rtol=1e-9, atol=1e-11)
            assert torch.allclose(ms.grad, ref.mean_scale.grad, rtol=1e-9, atol=1e-11)
        # the product backend is the CUDA library: CPU tensors are refused (no fallback)
        with pytest.raises(Exception):
            partitioned_graph_norm(x[lo:hi].float(), w.float(), b.float(), ms.float(), n)
        q.put((rank, "ok", None))
    except Exception:
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_partitioned_graph_norm_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gn_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"


# ----------------------------------------------------------------------------- the whole encoder over ranks
class _OracleBackedOps:
    """The three device operations of gmlm_b200.dist_encoder with oracle arithmetic (CPU, fp64): halo exchange
    with test-side row movers, the per-relation mean + transform of RGCNConvRef on the rank's rectangular edge
    list (root term on the local rows), GraphNorm+GELU over the partition with the emulated kernels."""

    def __init__(self, part, n_global):
        self.part, self.n_global = part, n_global

    def exchange(self, x_local):
        return halo_exchange(x_local, self.part, pack=_cpu_pack, unpack_add=_cpu_unpack_add)

    def conv(self, conv, X):
        from oracle import rgcn_propagate_mean_ref
        part = self.part
        w = conv.composed_weight()
        out = torch.zeros((part.n_local, conv.out_channels), dtype=X.dtype)
        for r in range(conv.num_relations):
            m = part.edge_type == r
            h = rgcn_propagate_mean_ref(X, part.edge_index[0, m], part.edge_index[1, m], part.n_local)
            out = out + h @ w[r]
        return out + X[: part.n_local] @ conv.root + conv.bias

    def norm_gelu(self, gn, y):
        from gmlm_b200.dist_norm import partitioned_graph_norm
        return partitioned_graph_norm(y, gn.weight, gn.bias, gn.mean_scale, self.n_global, gn.eps, True,
                                      backend=_EmulatedGraphNormKernels)


def _enc_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gmlm_b200.dist_encoder import PartitionedGraphEncoder, sync_gradients
        from gmlm_b200.partition import random_relabel
        from oracle import EncoderRef
        n, e, fin, hidden, out_dim = 240, 2600, 10, 4, 6
        ei = synth.rmat_edges(n, e, seed=21)
        ei, ranges, _ = random_relabel(ei, n, world)
        et = edge_type_bucket_ref(ei, n)                              # global out-degree, before partitioning
        g = torch.Generator().manual_seed(3)
        x = torch.randn(n, fin, dtype=torch.float64, generator=g)
        gout = torch.randn(n, out_dim, dtype=torch.float64, generator=g)
        torch.manual_seed(0)                                          # replicated parameters
        enc_full = EncoderRef(fin, hidden, out_dim, dropout_rate=0.0, use_checkpoint=False).double()
        with torch.no_grad():
            for k in range(1, 5):
                getattr(enc_full, f"gnorm{k}").mean_scale.uniform_(0.5, 1.0)
                getattr(enc_full, f"rgcn{k}").bias.uniform_(-0.1, 0.1)
        import copy
        enc_rank = copy.deepcopy(enc_full)
        # whole graph, one process
        x_full = x.clone().requires_grad_(True)
        fused_full = enc_full(x_full, ei, et)
        (fused_full * gout).sum().backward()
        # this rank's rows
        part = build_local_part(ei, et, ranges, rank)
        lo, hi = ranges[rank]
        model = PartitionedGraphEncoder(enc_rank, _OracleBackedOps(part, n))
        x_local = x[lo:hi].clone().requires_grad_(True)
        fused = model(x_local)
        (fused * gout[lo:hi]).sum().backward()
        sync_gradients(enc_rank)
        assert torch.allclose(fused, fused_full[lo:hi].detach(), rtol=1e-9, atol=1e-11)
        assert torch.allclose(x_local.grad, x_full.grad[lo:hi], rtol=1e-8, atol=1e-10)
        ref_grads = dict(enc_full.named_parameters())
        checked = 0
        for name, p in enc_rank.named_parameters():
            want = ref_grads[name].grad
            if want is None:
                assert p.grad is None, name                           # e.g. the dead residual_proj3
                continue
            assert p.grad is not None, name
            assert torch.allclose(p.grad, want, rtol=1e-7, atol=1e-9), (name, float((p.grad - want).abs().max()))
            checked += 1
        assert checked >= 4 * 7 + 4 + 9                               # convs, norms, residuals, fusion
        q.put((rank, "ok", None))
    except Exception:
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_encoder_equals_whole_graph_encoder_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_enc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"
