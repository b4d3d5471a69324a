"""CUDA bitstream (libb2c b2c_rans_*) vs the CPU oracle (oracle/rans_oracle.c): bytes bit-exact, round
trips, error flags, and the stream size against the reference's own bit estimate (model.py:71-74)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _freq(rng, C, L, prec, peak=4.0):
    A = 2 * L + 2
    p = rng.random((C, A)) + 0.05
    p[:, L] += peak
    p /= p.sum(1, keepdims=True)
    f = 1 + np.floor(p * (2 ** prec - A)).astype(np.int64)
    f[np.arange(C), f.argmax(1)] += 2 ** prec - f.sum(1)
    return f


def _w1_model(N=128):
    from golden_util import weights as W
    import iclr_17_compression_b200 as P
    m = P.ImageCompressor(N)
    m.load_state_dict(W(N, "W1"))
    return m.cuda().eval()


@pytest.mark.parametrize("shape,L,prec,bs", [((3, 8, 5, 7), 3, 12, 32), ((3, 8, 5, 7), 3, 12, 96), ((2, 16, 16, 16), 31, 14, 4096),
                                             ((1, 4, 3, 3), 1, 10, 32), ((2, 128, 4, 4), 31, 15, 512),
                                             ((2, 128, 32, 48), 31, 14, 32768), ((1, 192, 96, 128), 31, 14, 32768)])
def test_encode_bytes_match_oracle_and_round_trip(shape, L, prec, bs):
    from oracle import rans_oracle as R
    from iclr_17_compression_b200 import bitstream as BS
    rng = np.random.default_rng(sum(shape) + L)
    f = _freq(rng, shape[1], L, prec)
    z = np.rint(rng.normal(0, 2.0, shape)).astype(np.float32)
    z.reshape(-1)[0] = -300                                   # escapes
    z.reshape(-1)[-1] = 32767
    z.reshape(-1)[z.size // 2] = -32767
    tables = BS.RansTables.from_frequencies(f, prec, "cuda")
    assert np.array_equal(tables.frequencies(), f)
    want, off_w, bb_w, st_w = R.encode(z, R.pack_freq(f), L, prec, bs)
    out, off, bb, st = BS.encode_latents_raw(torch.from_numpy(z).cuda(), tables, bs)
    assert st.item() == 0 == st_w
    assert np.array_equal(bb.cpu().numpy().astype(np.uint32), bb_w)
    assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_w)
    got = out[:int(off[-1])].cpu().numpy()
    assert np.array_equal(got, want)                          # integer work: bit-exact
    # CUDA decoder on the CUDA stream, CUDA decoder on the oracle's stream, oracle decoder on the CUDA stream
    z2, st2 = BS.decode_latents_raw(out, off, tables, *shape, bs)
    assert st2.item() == 0 and np.array_equal(z2.cpu().numpy(), z)
    z3, st3 = R.decode(got, off.cpu().numpy(), R.pack_freq(f), L, prec, *shape, bs)
    assert st3 == 0 and np.array_equal(z3, z)
    # container: one byte string per image, and back
    streams = BS.encode_latents(torch.from_numpy(z).cuda(), tables, bs)
    assert len(streams) == shape[0] and sum(len(s) for s in streams) == len(want) + shape[0] * BS._HEAD.size + 4 * len(bb_w)
    assert np.array_equal(BS.decode_latents(streams, tables).cpu().numpy(), z)
    assert np.array_equal(BS.decode_latents(streams[-1:], tables).cpu().numpy(), z[-1:])


def test_tables_from_bit_estimator_match_fp64_rule():
    from oracle import rans_oracle as R
    from iclr_17_compression_b200 import bitstream as BS
    m = _w1_model()
    for L, prec in ((31, 14), (7, 10), (126, 15)):
        t = BS.RansTables.from_bit_estimator(m.bitEstimator, L, prec)
        f = t.frequencies()
        A, M = 2 * L + 2, 1 << prec
        assert f.shape == (128, A) and (f >= 1).all() and (f.sum(1) == M).all()
        f64 = R.frequencies64([p.detach().double().cpu().numpy().reshape(-1) for p in m.bitEstimator.param_list()], L, prec)
        d = np.abs(f - f64)
        # fp32 vs fp64 probabilities: a floor() may flip by one count per symbol; the remainder lands on the argmax
        rows = np.arange(128)
        best = f64.argmax(1)
        d_rest = d.copy()
        d_rest[rows, best] = 0
        assert d_rest.max() <= 1 and d[rows, best].max() <= A, (d_rest.max(), d[rows, best].max())
        # start / slot tables are consistent with the frequencies
        fc = t.freq_cum.cpu().numpy().view(np.uint32)
        start = np.concatenate([np.zeros((128, 1), np.int64), np.cumsum(f, 1)[:, :-1]], 1)
        assert np.array_equal(fc >> 16, start)
        ss = t.slot_sym.cpu().numpy()
        want = np.repeat(np.tile(np.arange(A, dtype=np.uint8), 128), f.reshape(-1)).reshape(128, M)
        assert np.array_equal(ss, want)


def test_errors_are_loud():
    from iclr_17_compression_b200 import bitstream as BS
    from iclr_17_compression_b200._lib import B2CError
    rng = np.random.default_rng(1)
    f = _freq(rng, 4, 3, 12)
    t = BS.RansTables.from_frequencies(f, 12, "cuda")
    z = torch.zeros(1, 4, 8, 8, device="cuda")
    z[0, 1, 2, 3] = 40000.0
    with pytest.raises(B2CError, match="outside"):
        BS.encode_latents(z, t, 64)
    z = torch.from_numpy(np.rint(rng.normal(0, 2, (1, 4, 8, 8))).astype(np.float32)).cuda()
    out, off, bb, st = BS.encode_latents_raw(z, t, 64, out_capacity=100)          # too small: flagged, nothing overrun
    assert st.item() & BS.ST_CAPACITY
    s = BS.encode_latents(z, t, 64)[0]
    bad = bytearray(s)
    bad[BS._HEAD.size + 4 * 4 + 2] ^= 0x10                                         # a bit of lane 0's state
    with pytest.raises(B2CError, match="corrupt"):
        BS.decode_latents([bytes(bad)], t)
    with pytest.raises(B2CError):
        BS.encode_latents(z, t, 48)                                                # block_symbols % 32
    with pytest.raises(B2CError):
        BS.encode_latents(z.cpu(), t, 64)                                          # no CPU path
    with pytest.raises(ValueError):
        BS.decode_latents([s], BS.RansTables.from_frequencies(_freq(rng, 4, 5, 12), 12, "cuda"))


@pytest.mark.parametrize("kind", ["U", "S"])
def test_compress_decompress_kodak_shape(kind):
    """cfg2 shape: decompress(compress(x)) is the model's own reconstruction bit for bit, the latents survive exactly,
    and the files are as large as the reference's estimate says (model.py:71-74) up to the table precision + states."""
    from oracle import balle17_oracle as O
    from iclr_17_compression_b200 import bitstream as BS
    m = _w1_model()
    x = O.synth_images(kind, 4, 512, 768, 1002 if kind == "U" else 1003).cuda()
    with torch.no_grad():
        clipped, z, bpp = m(x)
    tables = BS.RansTables.from_bit_estimator(m.bitEstimator)
    streams, image_bits = BS.compress(m, x, tables)
    assert len(streams) == 4
    assert torch.equal(BS.decode_latents(streams, tables), z)
    recon = BS.decompress(m, streams, tables)
    assert torch.equal(recon, clipped)
    est = image_bits.double().cpu().numpy()
    assert abs(est.sum() / (4 * 512 * 768) - bpp.item()) <= 1e-5 * bpp.item()
    from oracle import rans_oracle as R
    f = tables.frequencies()
    zc = z.cpu().numpy()
    for n, s in enumerate(streams):
        hd, bb, payload = BS.parse_stream(s)
        states = 128 * hd["nblk"] * 8
        bits = 8 * len(payload)
        # coder efficiency: within 1 % of what these tables allow, plus the 32 states per block
        ideal = R.ideal_bits(zc[n:n + 1], f, tables.half_range, tables.prec)
        assert ideal <= bits <= 1.01 * ideal + states + 64, (bits, ideal)
        # against the reference's estimate: f / 2^prec >= p * (1 - A / 2^prec), i.e. at most 0.0057 bit per latent more
        # than -log2(p), and an escaped value costs at most 16 bits more than its (smaller) point probability
        n_esc = int((np.abs(zc[n]) > tables.half_range).sum())
        assert bits <= 1.01 * (est[n] + 0.006 * zc[n].size + 16 * n_esc) + states + 64, (bits, est[n], n_esc)


def test_fp32_path_model_and_small_width():
    """ImageCompressor(32) (no tensor-core path): same round trip through the fp32 kernels"""
    from golden_util import Case
    from iclr_17_compression_b200 import bitstream as BS
    import iclr_17_compression_b200 as P
    c = Case("eval_W1_U_n32_64")
    m = P.ImageCompressor(c.N)
    m.load_state_dict(c.sd)
    m = m.cuda().eval()
    with torch.no_grad():
        clipped, z, bpp = m(c.x.cuda())
    streams, _ = BS.compress(m, c.x.cuda(), block_symbols=64)
    assert torch.equal(BS.decompress(m, streams), clipped)
    m.train()
    with pytest.raises(RuntimeError):
        BS.compress(m, c.x.cuda())


def test_codec_files_round_trip(tmp_path):
    """codec.py: image file -> .b2c file -> image file, through the command-line entry; the decoded image equals the
    model's own clipped reconstruction of the padded input (model.py:57-59), cropped and rounded to 8 bits"""
    Image = pytest.importorskip("PIL.Image")
    from iclr_17_compression_b200 import codec
    from oracle import resample_oracle as RS
    img = RS.synth_photo(50, 70, 3)
    Image.fromarray(img, "RGB").save(tmp_path / "a.png")
    codec.main(["encode", str(tmp_path / "a.png"), str(tmp_path / "a.b2c")])
    codec.main(["decode", str(tmp_path / "a.b2c"), str(tmp_path / "a_rec.png")])
    rec = np.asarray(Image.open(tmp_path / "a_rec.png"))
    assert rec.shape == (50, 70, 3)
    model = codec.build_model(128)
    x = torch.from_numpy(codec.pad_to_16(img)).cuda().permute(2, 0, 1).unsqueeze(0).float().div(255).contiguous()
    with torch.no_grad():
        ref, z, bpp = model(x)
    want = ref[0, :, :50, :70].mul(255).round().clamp(0, 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
    assert np.array_equal(rec, want)
    nbytes = (tmp_path / "a.b2c").stat().st_size
    est = bpp.item() * 64 * 80 / 8
    assert 0.95 * est < nbytes < est * 1.15 + 16 + 44 + 128 + 8  # estimate, + coder slack, headers, one state flush
    with pytest.raises(ValueError):
        codec.decode_image(codec.build_model(192), (tmp_path / "a.b2c").read_bytes())
