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


def _decode_for(wire, Q, R, K, l, i):
    """coding="for" (include/qppfusion.h): block b = l * ceil(K/32) + i/32 holds base[b] + 32 offsets of w_b bits."""
    from qppfusion import ops
    _o_ids, o_base, o_off, o_pay, _end = ops._for_segments(Q, R, K, wire.id_bits)
    buf = wire.words.numpy()
    base, off, pay = buf[o_base:o_off].view(np.int32), buf[o_off:o_pay].view(np.uint32), buf[o_pay:].view(np.uint32)
    b = l * ((K + 31) // 32) + (i >> 5)
    w = int(off[b + 1]) - int(off[b])
    v = 0
    if w:
        bit = (i & 31) * w
        wi, sh = bit >> 5, bit & 31
        x = int(pay[int(off[b]) + wi]) >> sh
        if sh + w > 32:
            x |= int(pay[int(off[b]) + wi + 1]) << (32 - sh)
        v = x & ((1 << w) - 1)
    idw = buf[:o_base].view(np.uint64)
    return _decode(idw, l * K + i, wire.id_bits, 0, 0)[0], int(np.int32(np.uint32((int(base[b]) + v) & 0xFFFFFFFF)))


def _sorted_arrays(seed, Q, R, K, max_id, scale):
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, max_id + 1, (Q, R, K)).astype(np.int32)
    sc = -np.sort(-(rng.beta(2, 5, (Q, R, K)) - 0.2) * scale, axis=2)
    micro = np.round(sc * 1e6).astype(np.int32)
    lens = rng.integers(0, K + 1, (Q, R)).astype(np.int32)
    lens[0, 0] = K
    micro[1 % Q, 0, :] = 777                               # constant list: zero-width blocks
    return torch.from_numpy(ids), torch.from_numpy(micro), torch.from_numpy(lens)


@pytest.mark.parametrize("shape,max_id,scale", [((9, 5, 1000), 2_681_467, 30.0), ((4, 3, 67), 1000, 1.0),
                                                ((3, 2, 31), 7, 2000.0), ((2, 1, 1), 0, 1.0)])
def test_frame_of_reference_packer_matches_the_documented_layout(shape, max_id, scale):
    from qppfusion import ops
    Q, R, K = shape
    ids, micro, lens = _sorted_arrays(5, Q, R, K, max_id, scale)
    w = ops.wire_pack(ids, micro, lens, coding="for", n_threads=3)
    assert w.coding == "for" and w.sc_bits == 0 and 0 < w.used <= w.words.numel()
    idf, mf, ln = ids.numpy().reshape(Q * R, K), micro.numpy().reshape(Q * R, K), lens.numpy().reshape(-1)
    rng = np.random.default_rng(6)
    for _ in range(min(Q * R * K, 15000)):
        l, i = int(rng.integers(0, Q * R)), int(rng.integers(0, K))
        if i < ln[l]:
            assert _decode_for(w, Q, R, K, l, i) == (int(idf[l, i]), int(mf[l, i])), (l, i)
    w1 = ops.wire_pack(ids, micro, lens, coding="for", n_threads=1)
    assert w1.used == w.used and torch.equal(w1.words[:w.used], w.words[:w.used])
    if K >= 1000:                                          # sorted lists: fewer bytes than the fixed-width stream
        assert w.used < ops.wire_pack(ids, micro, lens).words.numel()


def test_frame_of_reference_full_int32_range():
    from qppfusion import ops
    ids = torch.zeros((1, 1, 64), dtype=torch.int32)
    micro = torch.zeros((1, 1, 64), dtype=torch.int32)
    micro[0, 0, 0], micro[0, 0, 31], micro[0, 0, 40] = 2**31 - 1, -2**31, -5          # 32-bit wide block
    lens = torch.full((1, 1), 64, dtype=torch.int32)
    w = ops.wire_pack(ids, micro, lens, coding="for")
    for i in (0, 1, 31, 32, 40, 63):
        assert _decode_for(w, 1, 1, 64, 0, i)[1] == int(micro[0, 0, i])


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
@pytest.mark.parametrize("shape,max_id,scale", [((64, 5, 1000), 2_681_467, 30.0), ((4, 3, 67), 1000, 1.0),
                                                ((3, 2, 31), 7, 2000.0)])
def test_frame_of_reference_expansion_is_lossless(shape, max_id, scale):
    from qppfusion import ops
    Q, R, K = shape
    ids, micro, lens = _sorted_arrays(7, Q, R, K, max_id, scale)
    dev = torch.device("cuda:0")
    w = ops.wire_pack(ids, micro, lens, pin=True, coding="for").to(dev)
    d_ids = torch.full((Q, R, K), -9, dtype=torch.int32, device=dev)
    d_mi = torch.full((Q, R, K), -9, dtype=torch.int32, device=dev)
    ops.wire_unpack_(w, d_ids, d_mi)
    valid = (torch.arange(K)[None, None, :] < lens[:, :, None])
    assert torch.equal(d_ids.cpu()[valid], ids[valid]) and torch.equal(d_mi.cpu()[valid], micro[valid])


@pytest.mark.gpu
@pytest.mark.parametrize("coding", ["fixed", "for"])
def test_step_graph_fed_through_the_wire_equals_the_plain_step(coding):
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
    wire = ops.wire_pack(ids.cpu(), micro.cpu(), lens.cpu(), pin=True, coding=coding)
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
