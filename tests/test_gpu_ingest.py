"""Ingest / converters: pageable and pinned host memory, raw fp64 files — whatever the path,
the observations must land in HBM bit-exactly and in order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pageable_append_larger_than_the_staging_buffers(native, gpu_ctx):
    m, f = 9_000_001, 2           # X = 144 MB: several 64 MB staging chunks, ragged tail
    rng = np.random.default_rng(12345)
    X = rng.random((m, f)); y = rng.random(m)
    ds = native.Dataset(gpu_ctx, m, f)
    ds.append(X[:4_000_000], y[:4_000_000])
    ds.append(X[4_000_000:], y[4_000_000:])
    assert ds.size() == m
    for b, rows in ((0, 1000), (3_999_990, 20), (m - 7, 7)):
        Xr, yr = ds.read(b, rows)
        np.testing.assert_array_equal(Xr, X[b:b + rows]); np.testing.assert_array_equal(yr, y[b:b + rows])
    Xr, yr = ds.read(0, m)
    assert np.array_equal(Xr, X) and np.array_equal(yr, y)
    ds.free()


def test_pinned_append(native, gpu_ctx):
    m = 1_000_003
    hx, hy = native.PinnedBuffer(m), native.PinnedBuffer(m)
    hx.array[:] = np.linspace(0, 1, m); hy.array[:] = np.arange(m)
    ds = native.Dataset(gpu_ctx, m, 1)
    ds.append_raw(hx.array.ctypes.data, hy.array.ctypes.data, m)
    Xr, yr = ds.read(0, m)
    assert np.array_equal(Xr[:, 0], hx.array) and np.array_equal(yr, hy.array)
    ds.free(); hx.free(); hy.free()


def test_load_raw_files(native, oracle, gpu_ctx, tmp_path):
    fam = oracle.LINEAR
    m, f = 300_007, 5
    truth = np.arange(6, dtype=float)
    X, y = oracle.generate(fam, m, f, truth)
    px, py = tmp_path / "X.f64", tmp_path / "y.f64"
    X.tofile(px); y.tofile(py)
    ds = native.Dataset(gpu_ctx, m, f)
    ds.load_raw(str(px), str(py), 100_000)                       # first part
    ds.load_raw(str(px), str(py), m - 100_000, file_row_offset=100_000)
    assert ds.size() == m
    Xr, yr = ds.read(0, m)
    assert np.array_equal(Xr, X) and np.array_equal(yr, y)
    loss, grad = ds.loss_grad(fam, truth * 0.9)
    assert loss == pytest.approx(oracle.apply(fam, X, y, truth * 0.9), rel=1e-12)
    with pytest.raises(native.NativeError):                      # reading past the end of the file
        native.Dataset(gpu_ctx, 10, f).load_raw(str(px), str(py), 10, file_row_offset=m - 5)
    with pytest.raises(native.NativeError):
        native.Dataset(gpu_ctx, 10, f).load_raw(str(tmp_path / "missing"), str(py), 10)
    ds.free()


def test_a_view_of_a_freed_or_cleared_data_set_fails_loudly(native, oracle, gpu_ctx):
    """so_dataset_slice views alias the parent's memory (scalaopt_cuda.h): after the parent is cleared or freed, an evaluation
    on the view is SO_ESTATE, not a read of stale or released device memory."""
    fam = oracle.LINEAR
    X, y = oracle.generate(fam, 1000, 3, np.array([1.0, 2.0, 3.0, 4.0]))
    p = np.ones(4)
    ds = native.Dataset(gpu_ctx, 1000, 3).append(X, y)
    v = ds.slice(10, 500)
    vv = v.slice(5, 100)                      # a view of a view hangs on the owner
    loss, _ = v.loss_grad(fam, p)
    assert loss == pytest.approx(oracle.apply(fam, X[10:500], y[10:500], p), rel=1e-12)
    ds.clear()
    ds.append(X[:600], y[:600])               # re-ingested: the old views stay stale
    for view in (v, vv):
        with pytest.raises(native.NativeError) as e:
            view.loss_grad(fam, p)
        assert e.value.code == -6
    with pytest.raises(native.NativeError):
        v.read(0, 1)
    v2 = ds.slice(10, 500)                    # a fresh view works
    assert v2.loss_grad(fam, p)[0] == pytest.approx(loss, rel=1e-15)
    ds.free()
    with pytest.raises(native.NativeError) as e:
        v2.loss_grad(fam, p)
    assert e.value.code == -6
    for view in (v, vv, v2):
        view.free()
