"""Host logic of the graph cache (gmlm_b200/graph.py), on CPU: the device work (checksum kernel, CSR build) is
replaced by host stand-ins so that the two-level lookup, its eviction rules and its locking run here.  The same
cache with the real kernels behind it is covered by tests/test_gpu_graph.py."""
import threading

import pytest
import torch

import gmlm_b200.graph as gg


@pytest.fixture()
def host_cache(monkeypatch):
    built = []

    def fake_build(edge_index, edge_type, num_nodes, num_relations):
        built.append((edge_index, edge_type))
        return object()

    def fake_content_key(t):
        if t is None:
            return None
        return (tuple(t.shape), str(t.device), int(t.sum()), int((t * torch.arange(t.numel()).view_as(t)).sum()))

    monkeypatch.setattr(gg, "_require_cuda", lambda t, name: None)
    monkeypatch.setattr(gg, "_content_key", fake_content_key)
    monkeypatch.setattr(gg.RelGraph, "build", staticmethod(fake_build))
    gg.clear_graph_cache()
    yield built
    gg.clear_graph_cache()


def _graph(e=50, n=10, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g), torch.randint(0, 4, (e,), generator=g)


def test_identity_then_content_then_build(host_cache):
    ei, et = _graph()
    s0 = dict(gg.cache_stats)
    a = gg.get_rel_graph(ei, et, 10, 5)
    assert gg.get_rel_graph(ei, et, 10, 5) is a
    assert gg.cache_stats["identity_hits"] == s0["identity_hits"] + 1
    # a fresh tensor with the same content (what main.py:255 produces on every call): no rebuild
    assert gg.get_rel_graph(ei.clone(), et.clone(), 10, 5) is a
    assert gg.cache_stats["content_hits"] == s0["content_hits"] + 1 and len(host_cache) == 1
    # an in-place edit bumps _version: the identity key no longer matches and the content differs
    et[0] = (et[0] + 1) % 4
    assert gg.get_rel_graph(ei, et, 10, 5) is not a and len(host_cache) == 2
    # other arguments are part of both keys
    assert gg.get_rel_graph(ei, et, 11, 5) is not a and len(host_cache) == 3


def test_alias_table_is_bounded_in_bytes_and_keeps_the_newest(host_cache, monkeypatch):
    ei, et = _graph(e=1000)
    per_alias = ei.numel() * 8 + et.numel() * 8
    monkeypatch.setattr(gg, "_ALIAS_BYTES", 3 * per_alias)
    a = gg.get_rel_graph(ei, et, 10, 5)
    fresh = [et.clone() for _ in range(10)]                      # ten calls, each with a new edge_type tensor
    for t in fresh:
        assert gg.get_rel_graph(ei, t, 10, 5) is a
    assert len(host_cache) == 1                                  # never a rebuild
    assert len(gg._ALIAS) == 3 and sum(v[2] for v in gg._ALIAS.values()) <= 3 * per_alias
    assert gg.get_rel_graph(ei, fresh[-1], 10, 5) is a           # the newest alias is still an identity hit
    before = gg.cache_stats["content_hits"]
    assert gg.get_rel_graph(ei, fresh[0], 10, 5) is a            # an evicted alias costs a checksum, not a build
    assert gg.cache_stats["content_hits"] == before + 1 and len(host_cache) == 1
    # a single alias larger than the whole budget still stays (the current graph always hits by identity)
    monkeypatch.setattr(gg, "_ALIAS_BYTES", 1)
    gg.get_rel_graph(ei, et, 10, 5)
    assert len(gg._ALIAS) == 1
    hits = gg.cache_stats["identity_hits"]
    gg.get_rel_graph(ei, et, 10, 5)
    assert gg.cache_stats["identity_hits"] == hits + 1


def test_alias_count_bound_and_graph_lru(host_cache, monkeypatch):
    monkeypatch.setattr(gg, "_ALIAS_SIZE", 4)
    monkeypatch.setattr(gg, "_CACHE_SIZE", 2)
    graphs = [_graph(seed=s) for s in range(3)]
    objs = [gg.get_rel_graph(ei, et, 10, 5) for ei, et in graphs]
    assert len(gg._CACHE) == 2 and len(host_cache) == 3
    assert gg.get_rel_graph(*graphs[2], 10, 5) is objs[2]        # the two most recent graphs are kept ...
    assert gg.get_rel_graph(*graphs[1], 10, 5) is objs[1]
    assert gg.get_rel_graph(*graphs[0], 10, 5) is not objs[0]    # ... the oldest was dropped and is rebuilt
    for _ in range(10):
        gg.get_rel_graph(graphs[1][0], graphs[1][1].clone(), 10, 5)
    assert len(gg._ALIAS) <= 4


def test_concurrent_lookups_build_once(host_cache):
    ei, et = _graph(e=2000)
    out, errs = [], []

    def worker():
        try:
            for _ in range(50):
                out.append(gg.get_rel_graph(ei, et.clone(), 10, 5))
        except Exception as exc:                                 # noqa: BLE001
            errs.append(exc)

    threads = [threading.Thread(target=worker) for _ in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs and len(host_cache) == 1 and all(o is out[0] for o in out)
