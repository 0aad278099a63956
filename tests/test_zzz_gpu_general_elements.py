"""The hot path on general-element meshes (hexes, prisms, pyramids, tets, quadrilateral boundary faces:
tests/golden/elem_mixed, elem_pyramid -- boxmesh.mixed_box through the reference's UGRID reader) against the reference's
explicit iteration, bit for bit: the same phase tests tests/test_gpu_parity.py runs on the tetrahedral fixtures.  The
kernels are edge-based and see the element types only through the edge / half-edge lists (the prism mesh cube_LowFi has
been in the GPU lists since round 1); these fixtures were added after the round's GPU minutes were spent, so this file
sorts last.  The C oracle runs the same fixtures bit-exactly on the CPU (tests/test_oracle.py: GENERAL)."""
import pytest

from tests import test_gpu_parity as P
from tests.test_oracle import GENERAL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GENERAL)
def test_general_elements_phases(name):
    P.test_lsq_coefficients(name)
    P.test_update_bcs(name)
    P.test_gradient_limiter_residual_timestep(name)
    P.test_explicit_update(name)
    P.test_explicit_iterate_composite(name)
