"""GPU test of compute_mean_std(decode='gpu') (SURVEY §8a-S1 with §8f-1): the statistics pass with its JPEG files
decoded on the device.  Written after the round's last GPU run (host logic covered by tests/test_shims_cpu.py with
stand-in kernels; both kernels it calls are covered by the other GPU tests); kept in its own file, last in
collection order."""
import numpy as np
import pytest

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200.synth import synth_planes

pytestmark = pytest.mark.gpu


def test_compute_mean_std_gpu_decode_equals_host_decode(cuda, tmp_path):
    """compute_mean_std(paths, decode='gpu'): same decoded pixels as cv2.imread, exact integer sums -> identical
    float64 results, in normal and verification mode (compute_stats_experiments.py:8-24)."""
    import cv2
    from recursion_cellular_image_classification_b200 import compute_stats_experiments as cse
    planes = synth_planes(41, n=3, H=96, W=96)
    d = tmp_path / "exp0" / "Plate1"
    d.mkdir(parents=True)
    paths = []
    for i in range(planes.shape[0]):
        for ch in range(6):
            p = str(d / ("B%02d_s%d_w%d.jpeg" % (2 + i // 2, 1 + i % 2, ch + 1)))
            cv2.imwrite(p, planes[i, ch], [cv2.IMWRITE_JPEG_QUALITY, 95])
            paths.append(p)
    mh, sh = cse.compute_mean_std(paths)
    mg, sg = cse.compute_mean_std(paths, decode="gpu", chunk=7)          # ragged chunks
    np.testing.assert_array_equal(mh, mg)
    np.testing.assert_array_equal(sh, sg)
    decoded = np.stack([cv2.imread(p, cv2.IMREAD_GRAYSCALE) for p in paths]).reshape(3, 6, 96, 96)
    om, os_ = O.compute_mean_std_arrays(decoded)
    np.testing.assert_allclose(mg, om, rtol=1e-12)
    np.testing.assert_allclose(sg, os_, rtol=1e-10)
    vm, vs = cse.compute_mean_std(paths, mean=mg, std=sg, decode="gpu")
    np.testing.assert_allclose(vm, 0, atol=1e-9)
    np.testing.assert_allclose(vs, 1, rtol=1e-9)
