"""Error behaviour and robustness of the GPU entry points through the C ABI: argument checks map
to the documented codes (nothing aborts or throws across the boundary), unaligned record
slices work, contexts and calls can be mixed across threads."""
import threading

import numpy as np
import pytest

import ibu_b200 as ibu
from oracle import oracle_c as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ibu.GpuContext(0, chunk_records=1 << 18)
    yield c
    c.close()


def test_length_checks_mirror_header_validate(ctx):  # header.rs:179-184 bounds on the kernels
    d = ctx.malloc(4096)
    for bc, umi, exc in [(0, 12, ibu.InvalidBarcodeLength), (33, 12, ibu.InvalidBarcodeLength),
                         (16, 0, ibu.InvalidUmiLength), (16, 33, ibu.InvalidUmiLength)]:
        with pytest.raises(exc):
            ctx.validate_reduce_async(d, 10, bc, umi, d)
        with pytest.raises(exc):
            ctx.unpack_async(d, 10, bc, umi, d, d)
        with pytest.raises(exc):
            ctx.pack_async(d, d, 10, bc, umi, d)
    ctx.free(d)


def test_null_and_misaligned_arguments(ctx):
    d = ctx.malloc(1 << 16)
    with pytest.raises(ibu.ArgError):
        ctx.unpack_async(None, 10, 16, 12, d, d)
    with pytest.raises(ibu.ArgError):
        ctx.unpack_async(d, 10, 16, 12, d + 8, d)  # outputs must be 16-byte aligned
    with pytest.raises(ibu.ArgError):
        ctx.pack_async(d + 4, d, 10, 16, 12, d)
    with pytest.raises(ibu.ArgError):
        ctx.barcode_count(d + 8, 10)
    with pytest.raises(ibu.ArgError):
        ctx.barcode_count(d, 10, mode=5)
    with pytest.raises(ibu.ArgError):
        ctx.sort_records(d, 10, d)  # must not alias
    with pytest.raises(ibu.ArgError):
        ibu.GpuContext(99)
    ctx.unpack_async(None, 0, 16, 12, None, None)  # n = 0 needs no buffers
    ctx.synchronize()
    ctx.free(d)


@pytest.mark.parametrize("offset", [0, 1, 2, 3, 5])
@pytest.mark.parametrize("n", [0, 1, 3, 131, 100_003])
def test_validate_reduce_on_unaligned_record_slices(ctx, offset, n):
    """A slice of a record array is only 8-byte aligned; K1 peels to the next 32-byte boundary."""
    recs = oc.generate_records(0, n + offset, 16, 12, 1, 200_000, 77)
    d = ctx.malloc(max(recs.nbytes, 1))
    r = ctx.malloc(64)
    if recs.nbytes:
        ctx.h2d(d, recs)
    ctx.validate_reduce_async(d + 24 * offset, n, 16, 12, r)
    ctx.synchronize()
    assert ctx.read_result(r) == oc.reduce_records(recs[offset:], 16, 12)
    ctx.free(d), ctx.free(r)


def test_two_threads_share_a_context(ctx):
    """Blocking host-buffer calls serialise on the context's slots instead of corrupting them."""
    recs = oc.generate_records(0, 700_001, 16, 12, 1, 50_000, 5)
    want = oc.reduce_records(recs, 16, 12)
    ob, ou, _, _ = oc.unpack_records(recs, 16, 12, 0)
    out, errs = {}, []

    def run(kind):
        try:
            if kind == "reduce":
                out[kind] = ctx.process_host(recs, 16, 12)
            else:
                out[kind] = ctx.unpack_host(recs, 16, 12)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(k,)) for k in ("reduce", "unpack")]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    assert out["reduce"] == want
    assert np.array_equal(out["unpack"][0], ob) and np.array_equal(out["unpack"][1], ou)


def test_two_contexts_on_one_device():
    recs = oc.generate_records(0, 300_000, 16, 12, 0, 0, 6)
    a, b = ibu.GpuContext(0, chunk_records=1 << 16), ibu.GpuContext(0, chunk_records=1 << 17, n_slots=2)
    try:
        assert a.process_host(recs, 16, 12) == b.process_host(recs, 16, 12) == oc.reduce_records(recs, 16, 12)
    finally:
        a.close(), b.close()


def test_launch_counter_counts_kernels(ctx):
    d, r = ctx.malloc(24 * 1000), ctx.malloc(64)
    before = ibu.launch_count()
    ctx.generate_records_async(d, 0, 1000, 16, 12, 0, 0, 1)
    ctx.validate_reduce_async(d, 1000, 16, 12, r)
    ctx.synchronize()
    assert ibu.launch_count() - before == 3  # generator, K1, and the fold of K1's spread result blocks
    ctx.free(d), ctx.free(r)
