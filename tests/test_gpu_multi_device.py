"""The one-process multi-device path of ldsr_em_batch (csrc/ldsr_abi.cu, em_batch): groups sharded over
the devices of the context, one worker thread + stream per device, results scattered straight into the
caller's arrays.  Mirrors the reference's fan-out over folds/members (R/LDS_reconstruction.R:373-381):
the result must not depend on how the groups were dealt.  Needs >= 2 GPUs (`gpurun --gpus 2`)."""
import threading

import numpy as np
import pytest

from ldsr_b200 import _lib
from tests.test_gpu_parity import _np213, rand_theta0

pytestmark = pytest.mark.gpu


def _need2():
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")


def _job(n_folds=7, n_rest=9, seed=12):
    y, u, mu, inst = _np213()
    rng = np.random.default_rng(seed)
    held = [np.sort(rng.choice(inst, 11, replace=False)) for _ in range(n_folds)]
    fg = np.repeat(np.arange(n_folds), n_rest)
    th0 = rand_theta0(rng, 3, 3, n_folds * n_rest)
    return [dict(y=y, u=u, v=u)], np.zeros(n_folds, dtype=int), held, fg, th0


def test_poll_callback_runs_on_the_calling_thread_only():
    """include/ldsr_b200.h: worker threads never call back into the host language (the R shim's callback
    runs R_CheckUserInterrupt, which must stay on R's main thread)."""
    _need2()
    ser, gs, held, fg, th0 = _job()
    seen = []
    poll = lambda: seen.append(threading.get_ident()) or False
    a = _lib.em_batch(ser, gs, held, fg, th0, 300, 1e-5, n_devices=2, poll=poll)
    b = _lib.em_batch(ser, gs, held, fg, th0, 300, 1e-5, n_devices=1)
    assert set(seen) <= {threading.get_ident()}
    for k in ("theta", "lik", "iters", "best", "X", "Y", "V", "J"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_poll_callback_interrupts_a_sharded_call():
    _need2()
    ser, gs, held, fg, th0 = _job(n_folds=40, n_rest=64)
    calls = []
    with pytest.raises(_lib.LdsrError) as ei:
        _lib.em_batch(ser, gs, held, fg, th0, 100000, 0.0, n_devices=2, chunk_iters=50,
                      poll=lambda: calls.append(1) or True)
    assert ei.value.code == _lib.ERR_INTERRUPTED and len(calls) >= 1


def test_group_without_fits_on_two_devices():
    """A group that has no fit costs nothing; the sharded call must not hand a device an empty sub-batch
    (it used to fail with LDSR_ERR_ARG where the single-device call succeeds)."""
    _need2()
    ser, gs, held, fg, th0 = _job(n_folds=2, n_rest=5)
    fg = np.zeros(5, dtype=int)  # group 1 has no fits
    th0 = th0[:5]
    a = _lib.em_batch(ser, gs, held, fg, th0, 60, 1e-5, n_devices=1)
    b = _lib.em_batch(ser, gs, held, fg, th0, 60, 1e-5, n_devices=2)
    assert b["best"][1] == -1 and np.isnan(b["X"][b["traj_ptr"][1]:]).all()
    for k in ("theta", "lik", "iters", "best", "X", "Y", "V", "J"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_sharded_call_with_two_series_of_different_width_and_trace():
    """Rows of different width (p + q + 6 < theta_stride), trajectories of different length and the
    likelihood trace all go through the scatter."""
    _need2()
    y, u, mu, inst = _np213()
    rng = np.random.default_rng(3)
    T2 = 90
    u2 = rng.standard_normal((2, T2))
    y2 = 0.4 * np.cumsum(rng.standard_normal(T2)) * 0.1 + 0.2 * u2[0]
    y2[:30] = np.nan
    ser = [dict(y=y, u=u, v=u), dict(y=y2, u=u2, v=None, p=2, q=1)]
    gs = np.array([0, 1, 0, 1, 1])
    held = [inst[:5], np.array([40, 41]), inst[10:20], np.array([], dtype=int), np.array([70])]
    fg = np.repeat(np.arange(5), 6)
    th0 = rand_theta0(rng, 3, 3, 30)
    th0[:, 9:] = np.nan  # narrow rows: [A, B1, B2, C, D1, Q, R, mu1, V1] then unused
    for f in range(30):
        if gs[fg[f]] == 1:
            th0[f, :9] = np.concatenate([[rng.uniform()], rng.uniform(-1, 1, 2), [rng.uniform()], [0.0], [1, 1, 0, 1]])
        else:
            th0[f] = rand_theta0(rng, 3, 3, 1)[0]
    a = _lib.em_batch(ser, gs, held, fg, th0, 80, 1e-5, n_devices=1, want_liks=True)
    b = _lib.em_batch(ser, gs, held, fg, th0, 80, 1e-5, n_devices=2, want_liks=True)
    for k in ("theta", "lik", "iters", "status", "best", "liks", "X", "Y", "V", "J"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_kernel_choice_follows_the_whole_batch_not_the_shard():
    """1 400 fits are above the small-batch scan kernel's automatic range, 700 per device are inside it: the workers
    of the sharded call must take the kernel the whole batch takes, or the last bits of the results would
    depend on the number of devices."""
    _need2()
    ser, gs, held, fg, th0 = _job(n_folds=14, n_rest=100)
    a = _lib.em_batch(ser, gs, held, fg, th0, 150, 1e-5, n_devices=1)
    b = _lib.em_batch(ser, gs, held, fg, th0, 150, 1e-5, n_devices=2)
    for k in ("theta", "lik", "iters", "best", "X", "Y", "V", "J"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
