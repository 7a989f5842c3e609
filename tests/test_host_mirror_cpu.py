"""CPU-only checks of the host-side mirror (finitediff.jl_b200/api.py): argument handling, dispatch keys, matrix types —
everything that does not need a device.  Compute always goes through libfdjac_b200.so (no CPU path exists)."""
import numpy as np
import pytest
import torch

from _util import tridiagonal_coo


@pytest.fixture(scope="module")
def pkg():
    import _bootstrap
    return _bootstrap.load_package()


def test_public_surface_matches_reference_names(pkg):
    # src/jacobians.jl: JacobianCache, finite_difference_jacobian!, resize!; src/epsilons.jl: default_relstep, compute_epsilon
    for name in ("JacobianCache", "finite_difference_jacobian_", "resize_", "default_relstep", "compute_epsilon",
                 "SparseMatrixCSC", "BandedMatrix", "Tridiagonal"):
        assert hasattr(pkg, name)
    assert pkg.finite_difference_jacobian_b is pkg.finite_difference_jacobian_


def test_fdtype_validation(pkg):
    from finitediff_jl_b200 import api
    assert api._fdtype_code("forward") == 0 and api._fdtype_code(":central") == 1
    with pytest.raises(ValueError, match="Unrecognized fdtype"):      # epsilons.jl:159-167
        api._fdtype_code("backward")
    with pytest.raises(ValueError):
        api._fdtype_code("hcentral")


def test_tridiagonal_structural_nonzeros(pkg):
    T = pkg.Tridiagonal(6, buf=torch.zeros(16, dtype=torch.float64))
    rows, cols, slots = T.findstructralnz()
    r2, c2, s2 = tridiagonal_coo(6)
    assert (rows == r2).all() and (cols == c2).all() and (slots == s2).all()
    T.buf[:] = torch.arange(16, dtype=torch.float64)
    D = T.to_dense()
    assert D[1, 0] == 0 and D[0, 0] == 5 and D[0, 1] == 11      # [dl; d; du]


def test_dense_layout_checks(pkg):
    from finitediff_jl_b200 import api
    J = pkg.zeros_colmajor(3, 5, "cpu")
    assert J.shape == (3, 5) and J.stride() == (1, 3)
    assert api._dense_ld(J) == (3, 5, 3)
    with pytest.raises(ValueError, match="column-major"):
        api._dense_ld(torch.zeros(3, 5, dtype=torch.float64))
    # dense 0/1 prototype -> column-major structural nonzeros (jacobians.jl:473-488)
    r, c = api._findstructralnz_dense(np.array([[1, 1], [0, 1]]))
    assert r.tolist() == [1, 1, 2] and c.tolist() == [1, 2, 2]


def test_plan_keys_distinguish_patterns(pkg):
    from finitediff_jl_b200 import api
    a = torch.tensor([1, 2, 3], dtype=torch.int64)
    k1 = api._index_key(a)
    a[0] = 5                                                      # in-place edit bumps the tensor version
    assert api._index_key(a) != k1
    assert api._index_key(range(1, 4)) == ("r", 1, 4, 1)
    assert api._index_key(np.array([1, 2, 3])) == api._index_key([1, 2, 3])
    S = pkg.SparseMatrixCSC(2, 2, torch.tensor([1, 2, 3]), torch.tensor([1, 2]), None)
    assert api._has_sparsestruct(S) and not api._has_sparsestruct(torch.zeros(2, 2))
    with pytest.raises(TypeError, match="Int64"):
        api._index_ptr(torch.tensor([1, 2], dtype=torch.int32))


def test_cpu_inputs_rejected_loudly(pkg):
    # no CPU implementation of the path: CPU tensors never reach a fallback
    x = torch.zeros(4, dtype=torch.float64)
    with pytest.raises(TypeError):
        pkg.JacobianCache(x, "forward")
    with pytest.raises(TypeError):
        pkg.finite_difference_jacobian_(pkg.zeros_colmajor(4, 4, "cpu"), lambda a, b: None, x, "forward")


def test_identity_colorvec_detection(pkg):
    # the dense column branch (sparsity === nothing) is only the reference's behaviour for colorvec == 1:n
    # (jacobians.jl:547-557); the mirror refuses any other colorvec there, whatever container holds it
    import numpy as np
    import torch
    api = pkg.api
    assert api._is_identity_colorvec(None, 5)
    assert api._is_identity_colorvec(range(1, 6), 5)
    assert not api._is_identity_colorvec(range(1, 5), 5)
    assert api._is_identity_colorvec(np.arange(1, 6), 5)
    assert api._is_identity_colorvec([1, 2, 3, 4, 5], 5)
    assert not api._is_identity_colorvec(np.array([1, 2, 1, 2, 1]), 5)
    assert api._is_identity_colorvec(torch.arange(1, 6), 5)
    assert not api._is_identity_colorvec(torch.tensor([1, 2, 3, 4, 4]), 5)
    assert not api._is_identity_colorvec(torch.arange(1, 5), 5)
