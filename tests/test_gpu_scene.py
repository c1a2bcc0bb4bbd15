"""-m gpu: the reference's scene-level known-answer tests (tests/kats.py) run
through the C ABI on the CUDA backend, one ray per call like the reference API."""
import pytest

import kats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kat", kats.ALL, ids=lambda f: f.__name__)
def test_scene_kat_on_gpu(gpu, kat):
    kat(gpu)
