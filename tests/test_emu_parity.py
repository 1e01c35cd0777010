"""CPU suite: the host engine (setup tables, batcher, reader, Ogg layer) and the UNMODIFIED kernel
source compiled against the CUDA emulator (tests/emu), checked against the oracle.  Sized to run in
a few minutes; the full-size versions of the same cases are in test_gpu_parity.py."""
import pytest

import cases


@pytest.mark.parametrize("name,stride", [("1test", 1), ("2test", 31), ("3test", 37), ("issue6test", 41)])
def test_stage_parity(emu_ctx, name, stride):
    assert cases.stage_parity(emu_ctx, name, stride=stride) > 0


@pytest.mark.parametrize("name,stride", [("2test", 61), ("3test", 73)])
def test_stage_parity_truncated_packets(emu_ctx, name, stride):
    assert cases.stage_parity(emu_ctx, name, stride=stride, truncate=True) > 0


@pytest.mark.parametrize("l1_bits", [3, 6])
def test_stage_parity_narrow_huffman_tables(emu_lib_path, l1_bits):
    """Narrow first-level tables push most codewords through the second-level table and the longest
    ones through the sorted-array fallback of Codebook.DecodeScalar's restatement (k1_decode)."""
    from vorbispizza_b200 import Context
    ctx = Context(0, lib_path=emu_lib_path)
    try:
        ctx.set("l1_bits", l1_bits)
        assert cases.stage_parity(ctx, "3test", stride=53) > 0
        assert cases.stage_parity(ctx, "2test", stride=97, truncate=True) > 0
    finally:
        ctx.close()


@pytest.mark.parametrize("clip", [True, False])
def test_batch_pcm_1test(emu_ctx, clip):
    cases.batch_pcm_parity(emu_ctx, "1test", clip)


def test_reader_1test(emu_ctx):
    total, calls = cases.reader_parity(emu_ctx, "1test", lookahead=7)
    assert total == 17318


def test_reader_2test_small_windows(emu_ctx):
    total, _ = cases.reader_parity(emu_ctx, "2test", lookahead=100)
    assert total == 315790


def test_reader_planar_partial(emu_ctx):
    cases.reader_planar_and_partial(emu_ctx, "1test")


def test_seek_1test(emu_ctx):
    cases.seek_parity(emu_ctx, "1test", [0, 1, 1000, 5000, 17317, 17318, 17319, 20000, 4000], nread=600, lookahead=4)


def test_seek_3test_few(emu_ctx):
    cases.seek_parity(emu_ctx, "3test", [100000, 250, 287000], nread=1500, lookahead=4)


def test_synth_imdct_ola(emu_ctx):
    cases.synth_parity(emu_ctx, channels=2, n_streams=2, n_blocks=14)


def test_synth_generic_block_sizes(emu_ctx):
    cases.synth_parity(emu_ctx, channels=1, n_streams=1, n_blocks=8, lg0=8, lg1=9)
    cases.synth_parity(emu_ctx, channels=3, n_streams=1, n_blocks=6, lg0=9, lg1=10, clip=True)


def test_decode_files(emu_ctx):
    cases.decode_files_parity(emu_ctx, ["1test", "1test"])


def test_decode_files_ramped_groups(emu_lib_path):
    """The bulk pipeline with several groups in flight: small first groups (1, 2, 4 of a group of 8), the
    device page scan running one group ahead, three rotating batches; every copy of the file must come out
    identical, and identical with the host page scan."""
    import numpy as np
    from vorbispizza_b200 import Context, decode_files
    data = cases.load_file("1test")
    outs = []
    for gpu_scan in (1, 0):
        ctx = Context(0, lib_path=emu_lib_path)
        try:
            ctx.set("bulk_group", 8)
            ctx.set("gpu_scan", gpu_scan)
            pcm, counts = decode_files(ctx, [data] * 13, clip=True)
        finally:
            ctx.close()
        assert len(set(int(c) for c in counts)) == 1
        per = pcm.reshape(13, -1)
        assert all(np.array_equal(per[0].view(np.uint32), per[i].view(np.uint32)) for i in range(1, 13))
        outs.append(per[0].copy())
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_decode_files_groups_close_by_bytes(emu_lib_path):
    """A pipeline group closes when its container images reach the byte budget (long files: a music library does
    not fit 256 files into the 32-bit offsets of one batch): same PCM whatever the grouping."""
    import numpy as np
    from vorbispizza_b200 import Context, decode_files
    datas = [cases.load_file("1test"), cases.load_file("2test")[:60000], cases.load_file("1test")] * 2
    outs = []
    for budget in (0, 70000, 1):     # default (one group), two or three files per group, one file per group
        ctx = Context(0, lib_path=emu_lib_path)
        try:
            ctx.set("bulk_group", 256)
            ctx.set("bulk_group_bytes", budget)
            pcm, counts = decode_files(ctx, datas, clip=True)
        finally:
            ctx.close()
        outs.append((pcm.copy(), counts.copy()))
    for pcm, counts in outs[1:]:
        assert np.array_equal(counts, outs[0][1]) and np.array_equal(pcm.view(np.uint32), outs[0][0].view(np.uint32))


def test_excerpts_batch(emu_ctx):
    """BASELINE config 5 in small: random-access excerpts, every one like a fresh reader's SeekTo + read."""
    n = cases.excerpts_parity(emu_ctx, ["1test", "2test"], n_excerpts=6, nread=1500,
                              extra_positions=(0, 1, 1024, -1, -300, 10 ** 7))
    assert n >= 12


def test_bulk_calls_in_any_order(emu_lib_path):
    """vpz_decode_excerpts and vpz_decode_files share the context's pipeline batches: either may come first."""
    import numpy as np
    from vorbispizza_b200 import Context, decode_excerpts, decode_files
    data = cases.load_file("1test")
    ctx = Context(0, lib_path=emu_lib_path)
    try:
        pcm, _, got = decode_excerpts(ctx, [data], [0, 0], [100, 5000], [300, 300])
        assert list(got) == [300, 300]
        full, counts = decode_files(ctx, [data, data, data, data], clip=True)
        n = int(counts[0])
        assert np.array_equal(full[:n].view(np.uint32), full[n:2 * n].view(np.uint32))
        assert np.array_equal(pcm[:300].view(np.uint32), full[100:400].view(np.uint32))
    finally:
        ctx.close()


def test_decode_files_s16(emu_ctx):
    cases.decode_files_s16_parity(emu_ctx, ["1test"])


def test_empty_inputs(emu_ctx):
    """Empty batches are valid and produce nothing (no kernel launch, no error)."""
    import numpy as np
    from vorbispizza_b200 import Batch, decode_excerpts, decode_files
    pcm, counts = decode_files(emu_ctx, [])
    assert pcm.size == 0 and counts.size == 0
    data = cases.load_file("1test")
    pcm, offsets, got = decode_excerpts(emu_ctx, [data], np.zeros(0, np.uint32), np.zeros(0, np.int64), np.zeros(0, np.int32))
    assert pcm.size == 0 and offsets.size == 0 and got.size == 0
    # an excerpt of zero samples beside a normal one
    pcm, offsets, got = decode_excerpts(emu_ctx, [data], [0, 0], [100, 5000], [0, 50])
    assert pcm.size == 50 and list(got) == [0, 50] and list(offsets) == [0, 0]
    with Batch(emu_ctx) as b:
        b.decode(clip=True)
        assert b.total_floats == 0


# ---- every packet class through every kernel path ---------------------------------------------------
# The TestFiles are mono / stereo, residue 1 / 2, power-of-two VQ dimensions: left alone they only take
# the gather spectrum kernel and the 256 / 2048 IMDCT kernel.  "force_general" routes them through the
# general spectrum kernel (one CTA per packet), the generic IMDCT kernel and (2) the full symbol kernel.
@pytest.fixture(scope="module", params=[1, 2])
def general_ctx(emu_lib_path, request):
    from vorbispizza_b200 import Context
    ctx = Context(0, lib_path=emu_lib_path)
    ctx.set("force_general", request.param)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("name,stride", [("1test", 1), ("2test", 61), ("3test", 73), ("issue6test", 83)])
def test_general_path_stage_parity(general_ctx, name, stride):
    assert cases.stage_parity(general_ctx, name, stride=stride) > 0


def test_general_path_truncated(general_ctx):
    assert cases.stage_parity(general_ctx, "3test", stride=131, truncate=True) > 0


def test_general_path_pcm(general_ctx):
    cases.batch_pcm_parity(general_ctx, "1test", True)
    cases.decode_files_parity(general_ctx, ["1test"])
    cases.decode_files_s16_parity(general_ctx, ["1test"])


# ---- generated streams: every setup shape the TestFiles do not have (tests/synthvorbis.py) -------------
import synthvorbis  # noqa: E402


@pytest.mark.parametrize("shape", list(synthvorbis.SHAPES))
def test_generated_shape_stage_parity(emu_ctx, shape):
    """Integer stages, residue and spectrum of generated packets (and truncated copies): residue 0 / 1 / 2,
    1-8 channels, several coupling steps and submaps, floor 0, long codes, odd VQ dimensions, ..."""
    assert cases.synth_stage_parity(emu_ctx, shape, seed=11, n_packets=10) == 20


@pytest.mark.parametrize("shape", ["res012_3ch", "multi_submap", "floor0_mixed", "posts_beyond_block", "blocks_512_4096"])
def test_generated_shape_stream_parity(emu_ctx, shape):
    assert cases.synth_stream_parity(emu_ctx, shape, seed=5, n_packets=10, clip=True, eos_trim=100) > 0


def test_generated_mixed_batch(emu_ctx):
    """One bulk call over streams of different kernel classes (gather / general / full K1 paths, fast and
    generic IMDCT): every stream must come out as if decoded alone."""
    shapes = [(s, 7 + i) for i, s in enumerate(synthvorbis.SHAPES)]
    cases.synth_mixed_batch_parity(emu_ctx, shapes, n_packets=5, with_files=["1test"])


def test_65_post_floor_is_refused(emu_ctx):
    """Floor1.Posts holds 64 values (Floor1.cs:17): a 65-post floor faults in the reference; both sides refuse."""
    import oracle_binding as ob
    from vorbispizza_b200 import VpzError
    st = synthvorbis.make_stream(1, dict(channels=1, res_types=(1,), floor_posts=65), n_packets=2)
    with pytest.raises(ob.OracleError):
        ob.OracleStream(st["ogg"])
    with pytest.raises(VpzError):
        emu_ctx.create_setup(st["id"], st["setup"])
