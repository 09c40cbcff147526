"""Bit-packed posting transport (csrc/qf_wire.cu): the host packer against a plain-Python decoder of the documented
layout (CPU), and the device expansion against the original arrays (GPU)."""
import numpy as np
import pytest
import torch


def _decode(words, i, id_bits, sc_bits, sc_base):
    """posting i occupies bits [i*W, (i+1)*W) of the little-endian uint64 stream (include/qppfusion.h)."""
    W = id_bits + sc_bits
    bit = i * W
    wi, sh = bit >> 6, bit & 63
    v = int(words[wi]) >> sh
    if sh + W > 64:
        v |= int(words[wi + 1]) << (64 - sh)
    v &= (1 << W) - 1
    return v & ((1 << id_bits) - 1), (v >> id_bits) + sc_base


def _arrays(seed, Q, R, K, max_id, lo, hi, full=False):
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, max_id + 1, (Q, R, K)).astype(np.int32)
    micro = rng.integers(lo, hi + 1, (Q, R, K)).astype(np.int32)
    lens = np.full((Q, R), K, dtype=np.int32) if full else rng.integers(0, K + 1, (Q, R)).astype(np.int32)
    ids[0, 0, 0], micro[0, 0, 0], lens[0, 0] = max_id, hi, max(1, lens[0, 0])
    if K > 1:
        lens[0, 0] = max(2, lens[0, 0])
        micro[0, 0, 1] = lo
    return torch.from_numpy(ids), torch.from_numpy(micro), torch.from_numpy(lens)


@pytest.mark.parametrize("shape,max_id,lo,hi", [((7, 3, 50), 2_681_467, 0, 1_000_000), ((3, 2, 67), 5, -7, 9),
                                                 ((2, 1, 1), 0, 3, 3), ((5, 4, 33), 2**31 - 1, -2**31, 2**31 - 1),
                                                 ((260, 5, 100), 12345, -1_000_000, 1_000_000)])
def test_host_packer_matches_the_documented_layout(shape, max_id, lo, hi):
    from qppfusion import ops
    Q, R, K = shape
    ids, micro, lens = _arrays(1, Q, R, K, max_id, lo, hi)
    w = ops.wire_pack(ids, micro, lens, n_threads=3)
    assert w.id_bits == max(1, int(max_id).bit_length()) and w.sc_base == lo
    assert w.sc_bits == max(1, int(hi - lo).bit_length())
    words = w.words.numpy().view(np.uint64)
    idf, mf, ln = ids.numpy().reshape(-1), micro.numpy().reshape(-1), lens.numpy().reshape(-1)
    n = Q * R * K
    pick = np.arange(n) if n <= 20000 else np.random.default_rng(2).integers(0, n, 20000)
    for i in pick:
        i = int(i)
        d = _decode(words, i, w.id_bits, w.sc_bits, w.sc_base)
        assert d == ((int(idf[i]), int(mf[i])) if i % K < ln[i // K] else (0, w.sc_base)), i
    # the thread count must not change a bit
    w1 = ops.wire_pack(ids, micro, lens, n_threads=1)
    assert torch.equal(w.words, w1.words)


def test_packer_rejects_values_that_do_not_fit():
    from qppfusion import ops, _native as N
    ids, micro, lens = _arrays(3, 4, 2, 16, 1000, 0, 5000, full=True)
    with pytest.raises(N.NativeError, match="does not fit"):
        ops.wire_pack(ids, micro, lens, widths=(9, 13, 0))          # ids need 10 bits
    with pytest.raises(N.NativeError, match="does not fit"):
        ops.wire_pack(ids, micro, lens, widths=(10, 13, 1))         # a score below sc_base
    bad = ids.clone()
    bad[1, 1, 3] = -1
    with pytest.raises(N.NativeError, match="negative doc id"):
        ops.wire_pack(bad, micro, lens)
    lens2 = lens.clone()
    lens2[1, 1] = 3                                                 # ... unless it is padding
    ops.wire_pack(bad, micro, lens2)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,max_id,lo,hi", [((64, 5, 1000), 2_681_467, 0, 1_000_000), ((3, 2, 67), 5, -7, 9),
                                                 ((5, 4, 33), 2**31 - 1, -2**31, 2**31 - 1)])
def test_device_expansion_is_lossless(shape, max_id, lo, hi):
    from qppfusion import ops
    Q, R, K = shape
    ids, micro, lens = _arrays(4, Q, R, K, max_id, lo, hi)
    dev = torch.device("cuda:0")
    w = ops.wire_pack(ids, micro, lens, pin=True).to(dev)
    d_ids = torch.full((Q, R, K), -9, dtype=torch.int32, device=dev)
    d_mi = torch.full((Q, R, K), -9, dtype=torch.int32, device=dev)
    ops.wire_unpack_(w, d_ids, d_mi)
    valid = (torch.arange(K)[None, None, :] < lens[:, :, None])
    assert torch.equal(d_ids.cpu()[valid], ids[valid]) and torch.equal(d_mi.cpu()[valid], micro[valid])
    assert (d_ids.cpu()[~valid] == 0).all() and (d_mi.cpu()[~valid] == w.sc_base).all()


@pytest.mark.gpu
def test_step_graph_fed_through_the_wire_equals_the_plain_step():
    """FusionStepGraph(wire=...): the step that expands the packed upload first gives the same fused run, bit for bit."""
    import bench
    from qppfusion import ops
    from qppfusion.pipeline import FusionStepGraph
    dev = torch.device("cuda:0")
    ids, micro, lens, feats = bench.synth_device(256, 5, 1000, 11, dev)
    plain = FusionStepGraph(ids, micro, lens, feats, compact=True)
    plain.replay()
    torch.cuda.synchronize()
    ref = [t_.clone() for t_ in (plain.out.lens, plain.out.ids, plain.out.scores, plain.weights)]
    wire = ops.wire_pack(ids.cpu(), micro.cpu(), lens.cpu(), pin=True)
    ids2, micro2 = torch.zeros_like(ids), torch.zeros_like(micro)
    g = FusionStepGraph(ids2, micro2, lens, feats, compact=True, wire=wire.to(dev))
    ids2.zero_(); micro2.zero_()
    g.replay()
    torch.cuda.synchronize()
    n = ref[0].long()
    m = torch.arange(ref[1].shape[1], device=dev)[None, :] < n[:, None]
    assert torch.equal(g.out.lens, ref[0]) and torch.equal(g.out.ids[m], ref[1][m])
    assert torch.equal(g.out.scores[m].view(torch.int64), ref[2][m].view(torch.int64))
    assert torch.equal(g.weights.view(torch.int64), ref[3].view(torch.int64))
