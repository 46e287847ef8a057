"""Copy engines of the runtime (tfcuda_memcpy_h2d_async / tfcuda_wait_uploads / tfcuda_memcpy_d2h_async / tfcuda_copy_sync and their
tf.cuda_* bindings): the fast host<->device path of SURVEY.md §8(f1) for pipelines - uploads and downloads on their own streams.
Checks ordering (kernels queued after wait_uploads see the upload; a download sees the kernels queued before it), that sources stay
alive while a download is in flight, and the blocking bulk path next to the reference's per-element tf.tensor / .numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_async_round_trip_is_ordered_with_kernels(tf_cuda):
    tf = tf_cuda
    n = 1 << 20

    def prog():
        a = tf.input([-1], tf.float32)
        return a * 2.0 + 1.0
    double = tf.compile(prog)
    src = tf.cuda_pinned_array([n], "float32")
    dst = [tf.cuda_pinned_array([n], "float32") for _ in range(4)]
    dev = [tf.cuda_tensor(np.zeros(n, np.float32)) for _ in range(2)]
    rng = np.random.default_rng(0)
    want = []
    for it in range(4):
        data = rng.random(n, dtype=np.float32)
        tf.cuda_copy_sync()          # the pinned source is reused: the previous upload must have consumed it
        src[...] = data
        tf.cuda_upload_async(dev[it % 2], src)
        tf.cuda_wait_uploads()
        out = double(dev[it % 2])
        tf.cuda_download_async(out, dst[it])
        del out                      # the in-flight download keeps the tensor alive
        want.append(data * np.float32(2.0) + np.float32(1.0))
    tf.cuda_copy_sync()
    for it in range(4):
        assert np.array_equal(dst[it], want[it]), f"iteration {it}"


def test_bulk_and_per_element_paths_agree(tf_cuda):
    tf = tf_cuda
    rng = np.random.default_rng(1)
    for dtype in (np.float32, np.int32, np.uint32):
        a = (rng.random((37, 129)) * 1000).astype(dtype)
        fast = tf.cuda_tensor(a)
        slow = tf.tensor(a)
        assert np.array_equal(tf.cuda_numpy(fast), a) and np.array_equal(np.array(slow.numpy), a)
        assert np.array_equal(np.array(fast.numpy), a) and np.array_equal(tf.cuda_numpy(slow), a)
        assert fast.shape == slow.shape
