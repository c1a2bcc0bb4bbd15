"""-m gpu: the product's shading math against a witness that shares no source with it.

The C++ oracle compiles the same `csrc/shade/*.h` as the GPU, so GPU-vs-oracle image tests cannot see a transliteration
slip in those headers. Here `Material::eval` and `Material::scatter_importance` run on the device through
`crt_debug_material_eval` / `crt_debug_material_scatter` and are compared with `oracle/openpbr_ref.py` — a float64
restatement written separately, straight from the reference's `material/openpbr.rs` and `material/brdf.rs` — for every
material of samples/openpbr_showcase.usda and samples/veach_mis.usda plus a set that reaches the remaining branches
(anisotropy, thin film on metal and under a coat, coat colour, non-dispersive and inverted-IOR glass), over grids of view /
light directions on both sides of the surface and over stratified 4-D sample blocks.

Tolerance: the device computes in f32 with `powf` / `cosf` of CUDA's libm, the witness in f64: relative 2e-3 of the value
(plus 1e-6 absolute) — three orders of magnitude tighter than any sign or argument slip would produce.
(Tried on the host twin of this file, tests/test_shade_independent.py: flipping the sign of the `-1` in brdf.h's
`ggx_lambda_aniso` fails it at once.)"""
import pytest

import shade_cases as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(S.MATERIALS))
def test_eval_matches_the_independent_restatement(gpu, name):
    S.check_eval(name, gpu.debug_material_eval)


@pytest.mark.parametrize("name", sorted(S.MATERIALS))
def test_scatter_matches_the_independent_restatement(gpu, name):
    S.check_scatter(name, gpu.debug_material_scatter)
