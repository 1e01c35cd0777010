"""GPU parity tests proper (`-m gpu`, real B200): every case calls the sm_100a kernels through the
C ABI (vorbispizza_b200/libvpz.so) and compares with the CPU oracle on the same inputs.
Nothing here reads /root/reference; there is no fallback -- a missing library or GPU fails."""
import hashlib

import numpy as np
import pytest

import cases
import oracle_binding as ob
from conftest import FILES, load_file
from vorbispizza_b200 import Batch, SynthBatch, decode_files

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", FILES)
def test_stage_parity_every_packet(gpu_ctx, name):
    """BASELINE config 2: decoded codeword indices, floor Y, partition classes bit-exact on ALL packets."""
    n = cases.stage_parity(gpu_ctx, name, stride=1)
    assert n == {"1test": 25, "2test": 310, "3test": 366, "issue6test": 606}[name]


@pytest.mark.parametrize("name", ["2test", "3test", "issue6test"])
def test_stage_parity_truncated_packets(gpu_ctx, name):
    """quirk Q3 (end-of-packet behaviour) needs synthetic truncation: no fixture exercises it."""
    assert cases.stage_parity(gpu_ctx, name, stride=7, truncate=True) > 0


@pytest.mark.parametrize("l1_bits", [3, 6, 11])
def test_stage_parity_other_huffman_table_widths(l1_bits):
    """First-level widths that push codewords through the second-level table and the sorted-array
    fallback (narrow) or resolve nearly everything in one load (wide): same symbols, same bits."""
    from vorbispizza_b200 import Context
    ctx = Context(0)
    try:
        ctx.set("l1_bits", l1_bits)
        for name in ("3test", "issue6test"):
            assert cases.stage_parity(ctx, name, stride=7) > 0
        assert cases.stage_parity(ctx, "2test", stride=11, truncate=True) > 0
    finally:
        ctx.close()


@pytest.mark.parametrize("clip", [True, False])
@pytest.mark.parametrize("name", FILES)
def test_batch_pcm(gpu_ctx, name, clip):
    cases.batch_pcm_parity(gpu_ctx, name, clip)


@pytest.mark.parametrize("clip", [True, False])
@pytest.mark.parametrize("name", FILES)
def test_reader_read_samples(gpu_ctx, name, clip):
    total, _ = cases.reader_parity(gpu_ctx, name, clip=clip)
    assert total == {"1test": 17318, "2test": 315790, "3test": 288094, "issue6test": 548160}[name]


@pytest.mark.parametrize("lookahead", [1, 3, 50, 0])
def test_reader_window_sizes(gpu_ctx, lookahead):
    cases.reader_parity(gpu_ctx, "3test", lookahead=lookahead, chunk=4096)


@pytest.mark.parametrize("name", FILES)
def test_reader_planar_partial(gpu_ctx, name):
    cases.reader_planar_and_partial(gpu_ctx, name)


@pytest.mark.parametrize("name", FILES)
def test_seek(gpu_ctx, name):
    total = {"1test": 17318, "2test": 315790, "3test": 288094, "issue6test": 548160}[name]
    rng = np.random.default_rng(0x5EED0005)
    pos = [0, 1, 127, 128, 1023, 1024, total - 1, total, total + 1, total + 100000]
    pos += [int(x) for x in rng.integers(0, total, 40)]
    cases.seek_parity(gpu_ctx, name, pos, nread=4096, lookahead=40)


def test_synth_vs_oracle(gpu_ctx):
    cases.synth_parity(gpu_ctx, channels=2, n_streams=4, n_blocks=64)
    cases.synth_parity(gpu_ctx, channels=2, n_streams=2, n_blocks=32, clip=True)
    cases.synth_parity(gpu_ctx, channels=1, n_streams=2, n_blocks=32)


@pytest.mark.parametrize("lg0,lg1,ch", [(8, 8, 1), (8, 13, 1), (8, 10, 3), (9, 12, 2), (8, 11, 6), (11, 11, 2), (10, 12, 8)])
def test_synth_generic_block_sizes(gpu_ctx, lg0, lg1, ch):
    cases.synth_parity(gpu_ctx, channels=ch, n_streams=2, n_blocks=12, lg0=lg0, lg1=lg1)


def test_synth_full_size_properties(gpu_ctx):
    """BASELINE config 3 at full size (65,536 stereo blocks): size-independent properties.
    (1) the transform + window + OLA is linear, and scaling by a power of two is exact in fp32, so
        decode(4 X) == 4 decode(X) bit for bit; (2) a sample of streams equals the oracle."""
    n_streams, n_blocks, ch = 64, 1024, 2
    rng = np.random.default_rng(0x5EED0001)
    flags = cases.synth_flags(rng, n_streams, n_blocks)
    spectra = cases.synth_spectra(rng, flags, ch, 256, 2048)
    with SynthBatch(gpu_ctx, ch, 8, 11, flags, spectra) as b:
        b.decode(clip=False)
        a = b.read_all().copy()
        per_stream = [(b.run_offset(s), b.run_samples(s)) for s in range(n_streams)]
    with SynthBatch(gpu_ctx, ch, 8, 11, flags, spectra * np.float32(4.0)) as b:
        b.decode(clip=False)
        a4 = b.read_all()
    assert np.array_equal((a * np.float32(4.0)).view(np.uint32), a4.view(np.uint32))
    off = 0
    for s in range(n_streams):
        nfl = int(sum((2048 if f & 1 else 256) // 2 * ch for f in flags[s]))
        if s in (0, 17, 63):
            ref = cases.reference_imdct_ola(flags[s], spectra[off:off + nfl], ch, 256, 2048)
            o, n = per_stream[s]
            cases.assert_pcm_close(a[o:o + n * ch].reshape(-1, ch), ref, "config 3 stream %d" % s)
        off += nfl


def test_decode_files(gpu_ctx):
    cases.decode_files_parity(gpu_ctx, FILES + ["3test", "1test"], clip=True)
    cases.decode_files_parity(gpu_ctx, FILES, clip=False)


def test_replicated_streams_checksum(gpu_ctx):
    """BASELINE config 4 shape (replicated TestFiles, one batch): every replica must produce the
    same bytes as the first copy of its file, whose PCM is checked against the oracle."""
    reps = 64
    datas = [load_file(n) for n in FILES]
    pcm, counts = decode_files(gpu_ctx, datas * reps, clip=True)
    off = 0
    first = {}
    for i in range(len(FILES) * reps):
        name = FILES[i % len(FILES)]
        s = ob.OracleStream(datas[i % len(FILES)])
        n = int(counts[i]) * s.channels
        h = hashlib.sha256(pcm[off:off + n].tobytes()).hexdigest()
        if name not in first:
            first[name] = h
            ref, _, _ = s.decode_all()
            cases.assert_pcm_close(pcm[off:off + n].reshape(-1, s.channels), ref, name)
        assert h == first[name], "replica %d of %s differs" % (i // len(FILES), name)
        off += n
    assert off == pcm.size


def test_offset_starts_match_seek_semantics(gpu_ctx):
    """Runs that start in the middle of a stream (config 4 'offset' replicas / config 5 excerpts):
    a run started at packet k equals the oracle's output from the sample where packet k+1 begins."""
    data = load_file("3test")
    s = ob.OracleStream(data)
    st = gpu_ctx.create_setup(s.header_packet(0), s.header_packet(2))
    pk = s.audio_packets()
    full, _, _ = s.decode_all()
    with Batch(gpu_ctx) as b:
        whole = b.add_run(st, [p["data"] for p in pk[:-1]])
        cnt = b.run_packet_samples(whole, len(pk) - 1)
        starts = [1, 7, 100, 200, 333]
        runs = [b.add_run(st, [p["data"] for p in pk[k:k + 20]]) for k in starts]
        b.decode(clip=True)
        for k, r in zip(starts, runs):
            got = b.read_run(r)
            begin = int(cnt[:k + 1].sum())
            cases.assert_pcm_close(got, full[begin:begin + got.shape[0]], "run from packet %d" % k)
    gpu_ctx.release_setup(st)


def test_excerpts_batch(gpu_ctx):
    """BASELINE config 5: random-access excerpts (SeekTo + 4,096 samples) in one bulk call vs the oracle's
    reader, incl. starts at 0 / 1 / block boundaries / the last samples / beyond the end."""
    n = cases.excerpts_parity(gpu_ctx, ["2test", "3test", "issue6test"], n_excerpts=300, nread=4096,
                              extra_positions=(0, 1, 127, 128, 1023, 1024, -1, -2, -4096, -5000, 10 ** 7))
    assert n >= 300


def test_excerpts_batch_unclipped_small_reads(gpu_ctx):
    cases.excerpts_parity(gpu_ctx, ["1test", "3test"], n_excerpts=64, nread=700, clip=False, seed=77)


@pytest.mark.parametrize("clip", [True, False])
def test_decode_files_s16(gpu_ctx, clip):
    """16-bit output fused into the IMDCT kernel (half the device-to-host bytes)."""
    cases.decode_files_s16_parity(gpu_ctx, FILES, clip=clip)


# ---- every kernel path under oracle parity -----------------------------------------------------------
# (1) the TestFiles forced through the general spectrum kernel, the generic IMDCT kernel and the full
#     symbol kernel ("force_general"), (2) generated streams of every setup shape the TestFiles do not have
#     (tests/synthvorbis.py): >= 1,000 packet decodes per shape, bit-exact integer stages / residue / spectrum.
import synthvorbis  # noqa: E402


@pytest.fixture(scope="module", params=[1, 2])
def general_ctx(request):
    from vorbispizza_b200 import Context
    ctx = Context(0)
    ctx.set("force_general", request.param)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("name", FILES)
def test_general_path_stage_parity_every_packet(general_ctx, name):
    n = cases.stage_parity(general_ctx, name, stride=1)
    assert n == {"1test": 25, "2test": 310, "3test": 366, "issue6test": 606}[name]


def test_general_path_truncated_packets(general_ctx):
    assert cases.stage_parity(general_ctx, "3test", stride=5, truncate=True) > 0
    assert cases.stage_parity(general_ctx, "2test", stride=7, truncate=True) > 0


@pytest.mark.parametrize("name", FILES)
def test_general_path_pcm(general_ctx, name):
    cases.batch_pcm_parity(general_ctx, name, True)
    cases.batch_pcm_parity(general_ctx, name, False)


def test_general_path_bulk_and_s16(general_ctx):
    cases.decode_files_parity(general_ctx, FILES, clip=True)
    cases.decode_files_s16_parity(general_ctx, FILES, clip=True)
    cases.reader_parity(general_ctx, "3test", lookahead=50, chunk=4096)


@pytest.mark.parametrize("shape", list(synthvorbis.SHAPES))
def test_generated_shape_stage_parity(gpu_ctx, shape):
    n = cases.synth_stage_parity(gpu_ctx, shape, seed=101, n_packets=260, mean_len=150)
    n += cases.synth_stage_parity(gpu_ctx, shape, seed=202, n_packets=260, mean_len=60)
    assert n >= 1000


@pytest.mark.parametrize("clip", [True, False])
@pytest.mark.parametrize("shape", list(synthvorbis.SHAPES))
def test_generated_shape_stream_parity(gpu_ctx, shape, clip):
    assert cases.synth_stream_parity(gpu_ctx, shape, seed=31 + int(clip), n_packets=120, clip=clip, lookahead=17,
                                     eos_trim=77) > 0


def test_generated_mixed_batch(gpu_ctx):
    """Streams of every kernel class in ONE bulk call beside the TestFiles: per-setup path selection."""
    shapes = [(s, 7 + i) for i, s in enumerate(synthvorbis.SHAPES)]
    cases.synth_mixed_batch_parity(gpu_ctx, shapes, n_packets=60, with_files=FILES, clip=False)
    cases.synth_mixed_batch_parity(gpu_ctx, shapes[::3], n_packets=30, with_files=["3test"], clip=True)


def test_mixed_batch_keeps_fast_streams_fast(gpu_ctx):
    """A batch of stereo TestFile streams plus ONE 6-channel stream: the stereo packets must still take the
    gather / fast kernels (same kernel time as without the odd stream, within noise)."""
    import time
    from vorbispizza_b200 import decode_files
    datas = [load_file("3test")] * 256
    odd = synthvorbis.make_stream(9, "ch6_coupled", n_packets=40)["ogg"]

    def best(ds):
        t = []
        for _ in range(4):
            t0 = time.perf_counter()
            decode_files(gpu_ctx, ds, clip=True)
            t.append(time.perf_counter() - t0)
        return min(t)
    base = best(datas)
    mixed = best(datas + [odd])
    # the batch-wide switch this replaces made such a batch 2-3x slower
    assert mixed < 1.5 * base + 0.02, (base, mixed)


def test_65_post_floor_is_refused(gpu_ctx):
    from vorbispizza_b200 import VpzError
    st = synthvorbis.make_stream(1, dict(channels=1, res_types=(1,), floor_posts=65), n_packets=2)
    with pytest.raises(VpzError):
        gpu_ctx.create_setup(st["id"], st["setup"])
