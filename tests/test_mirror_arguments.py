"""Argument validation of the API mirror that happens before any device work (reference:
src/pybindings/stream/model.rs:495-559,645-708,985-1060; pybindings/stream/stack.rs:217-241)."""
import numpy as np
import pytest


def test_model_argument_errors():
    from constriction_b200.stream import model as M
    with pytest.raises(ValueError):
        M.QuantizedGaussian(-5, 5, 0.0, 0.0)          # std must be positive
    with pytest.raises(ValueError):
        M.QuantizedGaussian(-5, 5, 0.0, -1.0)
    with pytest.raises(ValueError):
        M.Categorical(lazy=True, perfect=True)
    with pytest.raises(ValueError):
        M.QuantizedLaplace(-5, 5, 0.0, 0.0)           # scale must be positive
    with pytest.raises(ValueError):
        M.QuantizedCauchy(-5, 5, 0.0, -2.0)
    with pytest.raises(ValueError):
        M.CustomModel(lambda x: 0.5, lambda q: 0.0, 3, 3)
    assert M.Categorical()._perfect and M.Bernoulli()._perfect   # the reference's defaults (model.rs:508-523,1010-1050)
    assert not M.Categorical(lazy=True)._perfect and not M.Categorical(perfect=False)._perfect
    b = M.Binomial()
    with pytest.raises(TypeError):
        b._family_len((np.array([3.0]), np.array([0.5])))       # n must be int32
    assert b._family_len((np.array([3, 4], dtype=np.int32), np.array([0.5, 0.1]))) == 2
    fam = M.QuantizedGaussian(-5, 5)
    with pytest.raises(ValueError):
        fam._concrete_table()                          # "No model parameters specified."
    with pytest.raises(ValueError):
        fam._family_len((np.zeros(3),))                # wrong number of parameters
    with pytest.raises(ValueError):
        fam._family_len((np.zeros(3), np.ones(4)))     # unequal shapes
    with pytest.raises(TypeError):
        fam._family_len((np.zeros(3, dtype=np.int32), np.ones(3)))
    with pytest.raises(ValueError):
        fam._family_table((np.zeros(3), np.array([1.0, 0.0, 2.0])))   # a std that is not positive
    with pytest.raises(ValueError):
        M.QuantizedGaussian(-5, 5, 0.0, 1.0)._family_len((np.zeros(3),))   # already fully parameterised
    u = M.Uniform()
    with pytest.raises(TypeError):
        u._family_len((np.array([2.0, 3.0]),))         # sizes must be int32
    with pytest.raises(ValueError):
        u._family_table((np.array([1, 5], dtype=np.int32),))
    assert u._family_len((np.array([2, 5, 9], dtype=np.int32),)) == 3


def test_coder_argument_errors():
    from constriction_b200.stream import stack
    with pytest.raises(ValueError):
        stack.AnsCoder(seal=True)                      # "Need compressed data to seal."
    with pytest.raises(TypeError):
        stack.AnsCoder(np.array([1, 2, 3], dtype=np.int64))
    with pytest.raises(ValueError):
        stack.AnsCoder(np.array([5, 0], dtype=np.uint32))   # ANS data never ends in a zero word
    c = stack.AnsCoder(np.array([7, 9], dtype=np.uint32))
    assert c.pos() == (0, (9 << 32) | 7) and c.num_words() == 2 and not c.is_empty()
    assert np.array_equal(c.get_compressed(), [7, 9])
    with pytest.raises(ValueError):
        c.seek(1, 123)                                 # past the end of the bulk
