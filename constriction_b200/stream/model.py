"""Entropy models with the reference's Python names and semantics
(reference: src/pybindings/stream/model.rs, model/internals.rs).

A model object is either *concrete* (all parameters given: one CDF table) or a *family* (some
parameters deferred to per-symbol arrays passed to `encode*` / `decode`: one CDF row per symbol,
tabulated on the device in one launch, then addressed by a per-symbol model index)."""
from __future__ import annotations

import os

import numpy as np

from .. import batch as B


class Model:
    _table = None      # ModelTable of a concrete model (built lazily)
    _nparams = 0

    def _concrete_table(self) -> B.ModelTable:
        raise ValueError("No model parameters specified.")

    def _family_table(self, params) -> B.ModelTable:
        raise ValueError("Model parameters were specified but the model is already fully parameterized.")

    def _family_len(self, params) -> int:
        raise ValueError("Model parameters were specified but the model is already fully parameterized.")


def _float_param(p) -> np.ndarray:
    p = np.asarray(p)
    if p.ndim != 1 or p.dtype not in (np.float32, np.float64):
        raise TypeError("model parameters must be rank-1 numpy arrays with dtype float32 or float64")
    return p.astype(np.float64)  # internals.rs:169-174: f32 parameters are widened to f64


class QuantizedGaussian(Model):
    """pybindings/stream/model.rs:645-708; quantize.rs:284-308,525-568."""

    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, mean=None, std=None):
        self._lo, self._hi = int(min_symbol_inclusive), int(max_symbol_inclusive)
        self._mean, self._std = mean, std
        self._nparams = (mean is None) + (std is None)
        if self._nparams == 0 and not float(std) > 0.0:
            raise ValueError("Invalid model parameter: `std` must be positive.")

    def _concrete_table(self):
        if self._nparams:
            raise ValueError("No model parameters specified.")
        if self._table is None:
            self._table = B.ModelTable.quantized_gaussian(self._lo, self._hi, [float(self._mean)], [float(self._std)])
        return self._table

    def _split(self, params):
        if self._nparams == 0:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != self._nparams:
            raise ValueError(f"Wrong number of model parameters: expected {self._nparams}, got {len(params)}.")
        cols = [_float_param(p) for p in params]
        n = cols[0].size
        if any(c.size != n for c in cols):
            raise ValueError("Model parameters have unequal shape")
        it = iter(cols)
        means = next(it) if self._mean is None else np.full(n, float(self._mean))
        stds = next(it) if self._std is None else np.full(n, float(self._std))
        return means, stds

    def _family_len(self, params):
        return self._split(params)[0].size

    def _family_table(self, params):
        means, stds = self._split(params)
        if not np.all(stds > 0.0):
            raise ValueError("Invalid model parameter: `std` must be positive.")
        if os.environ.get("CTR_GAUSS_TABLES") == "1":  # one tabulated CDF row per symbol (tests compare both paths)
            return B.ModelTable.quantized_gaussian(self._lo, self._hi, means, stds)
        # default: parameters go to the device as they are and the kernels evaluate them (no table per symbol)
        return B.GaussianParams(self._lo, self._hi, means, stds)


class Categorical(Model):
    """pybindings/stream/model.rs:455-560.  perfect=False: fast / lazy quantisation (categorical.rs:16-54,
    lazy_contiguous.rs:131-167,228-330; both give the same table); perfect=True -- the reference's default when
    neither flag is given -- the optimal quantisation of categorical.rs:56-177."""

    def __init__(self, probabilities=None, lazy=None, perfect=None):
        if lazy is None and perfect is None:
            lazy, perfect = False, True
        elif lazy and perfect:
            raise ValueError("Both arguments `lazy` and `perfect` cannot be set to `True` at the same time.")
        else:
            lazy, perfect = bool(lazy), bool(perfect)
        self._perfect = perfect
        self._probs = None
        if probabilities is not None:
            p = np.asarray(probabilities)
            if p.ndim != 1 or p.dtype not in (np.float32, np.float64):
                raise TypeError("probabilities must be a rank-1 numpy array with dtype float32 or float64")
            self._probs = np.ascontiguousarray(p)
            self._table = self._make(self._probs)

    def _make(self, p):
        try:
            # perfect=True (the reference's default): categorical.rs:56-177; else fast_quantized_cdf
            return B.ModelTable.categorical_perfect(p) if self._perfect else B.ModelTable.categorical(p)
        except ValueError:
            raise ValueError("Probability distribution not normalizable (the array of probabilities\n"
                             "might be empty, contain negative values or NaNs, or sum to infinity).") from None

    def _concrete_table(self):
        if self._probs is None:
            raise ValueError("No model parameters specified.")
        return self._table

    def _matrix(self, params):
        if self._probs is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != 1:
            raise ValueError(f"Wrong number of model parameters: expected 1, got {len(params)}.")
        p = np.asarray(params[0])
        if p.ndim != 2 or p.dtype not in (np.float32, np.float64):
            raise TypeError("probabilities must be a rank-2 numpy array with dtype float32 or float64")
        return np.ascontiguousarray(p)

    def _family_len(self, params):
        return self._matrix(params).shape[0]

    def _family_table(self, params):
        return self._make(self._matrix(params))


class Uniform(Model):
    """pybindings/stream/model.rs:570-600; uniform.rs:44-146: every bin 2^24 // size, the last one takes the
    remainder.  `size` fixed, or one int32 `size` per symbol: the rows are then padded to the largest alphabet with
    zero-probability symbols, which can neither be encoded (KeyError, like a symbol >= size in the reference) nor
    come out of the decoder."""

    def __init__(self, size=None):
        self._size = None if size is None else int(size)
        self._nparams = 1 if size is None else 0

    def _concrete_table(self):
        if self._size is None:
            raise ValueError("No model parameters specified.")
        if self._table is None:
            self._table = B.ModelTable.uniform(self._size)
        return self._table

    def _sizes(self, params):
        if self._size is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != 1:
            raise ValueError(f"Wrong number of model parameters: expected 1, got {len(params)}.")
        sizes = np.asarray(params[0])
        if sizes.ndim != 1 or sizes.dtype != np.int32:
            raise TypeError("size must be a rank-1 numpy array with dtype int32")
        return sizes

    def _family_len(self, params):
        return self._sizes(params).size

    def _family_table(self, params):
        sizes = self._sizes(params).astype(np.int64)
        if sizes.size == 0:
            return B.ModelTable.uniform(2)
        if sizes.min() < 2 or sizes.max() > (1 << 24):
            raise ValueError("Invalid model parameter: `size` must be at least 2 and at most 2^24.")
        widest = int(sizes.max())
        if widest * sizes.size > (1 << 28):
            raise NotImplementedError("Uniform with per-symbol sizes: table too large (alphabet x symbols > 2^28)")
        per_bin = (1 << 24) // sizes
        i = np.arange(widest + 1, dtype=np.int64)[None, :]
        rows = np.where(i < sizes[:, None], i * per_bin[:, None], 1 << 24).astype(np.uint32)
        return B.ModelTable.from_cdf(rows)


class Bernoulli(Model):
    """pybindings/stream/model.rs:985-1060: a two-symbol categorical model over [1 - p, p] in f64, quantised by
    `fast_quantized_cdf` (perfect=False, categorical.rs:16-54) or -- the reference's default -- perfectly
    (categorical.rs:56-177); `p` fixed, or one `p` per symbol."""

    def __init__(self, p=None, perfect=None):
        self._perfect = True if perfect is None else bool(perfect)
        self._p = None if p is None else float(p)
        self._nparams = 1 if p is None else 0
        if self._p is not None:
            self._table = self._make(np.array([[1.0 - self._p, self._p]], dtype=np.float64))

    def _make(self, pmf):
        try:
            return B.ModelTable.categorical_perfect(pmf) if self._perfect else B.ModelTable.categorical(pmf)
        except ValueError:
            raise ValueError("`p` must be >= 0.0 and <= 1.0.") from None

    def _concrete_table(self):
        if self._p is None:
            raise ValueError("No model parameters specified.")
        return self._table

    def _params(self, params):
        if self._p is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != 1:
            raise ValueError(f"Wrong number of model parameters: expected 1, got {len(params)}.")
        return _float_param(params[0])

    def _family_len(self, params):
        return self._params(params).size

    def _family_table(self, params):
        p = self._params(params)
        return self._make(np.ascontiguousarray(np.stack([1.0 - p, p], axis=1)))


class _TwoParameterQuantized(Model):
    """Leaky quantisation of a two-parameter distribution, any subset of the parameters deferred to per-symbol arrays."""
    _kind = None
    _second = "scale"

    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, first=None, second=None):
        self._lo, self._hi = int(min_symbol_inclusive), int(max_symbol_inclusive)
        self._p0, self._p1 = first, second
        self._nparams = (first is None) + (second is None)
        if second is not None and not float(second) > 0.0:
            raise ValueError(f"Invalid model parameter: `{self._second}` must be positive.")

    def _concrete_table(self):
        if self._nparams:
            raise ValueError("No model parameters specified.")
        if self._table is None:
            self._table = B.ModelTable.quantized(self._kind, self._lo, self._hi, [float(self._p0)], [float(self._p1)])
        return self._table

    def _split(self, params):
        if self._nparams == 0:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != self._nparams:
            raise ValueError(f"Wrong number of model parameters: expected {self._nparams}, got {len(params)}.")
        cols = [_float_param(p) for p in params]
        n = cols[0].size
        if any(c.size != n for c in cols):
            raise ValueError("Model parameters have unequal shape")
        it = iter(cols)
        p0 = next(it) if self._p0 is None else np.full(n, float(self._p0))
        p1 = next(it) if self._p1 is None else np.full(n, float(self._p1))
        return p0, p1

    def _family_len(self, params):
        return self._split(params)[0].size

    def _family_table(self, params):
        p0, p1 = self._split(params)
        if not np.all(p1 > 0.0):
            raise ValueError(f"Invalid model parameter: `{self._second}` must be positive.")
        return B.ModelTable.quantized(self._kind, self._lo, self._hi, p0, p1)  # one CDF row per symbol


class QuantizedLaplace(_TwoParameterQuantized):
    """pybindings/stream/model.rs:740-830: QuantizedLaplace(min, max, mean=None, scale=None)."""
    _kind = "laplace"

    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, mean=None, scale=None):
        super().__init__(min_symbol_inclusive, max_symbol_inclusive, mean, scale)


class QuantizedCauchy(_TwoParameterQuantized):
    """pybindings/stream/model.rs:840-920: QuantizedCauchy(min, max, loc=None, scale=None)."""
    _kind = "cauchy"

    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, loc=None, scale=None):
        super().__init__(min_symbol_inclusive, max_symbol_inclusive, loc, scale)


class Binomial(Model):
    """pybindings/stream/model.rs:925-960: Binomial(n=None, p=None) over {0..n}; missing parameters come per symbol
    (`n` as int32, `p` as float arrays, in this order)."""

    def __init__(self, n=None, p=None):
        self._n = None if n is None else int(n)
        self._p = None if p is None else float(p)
        self._nparams = (n is None) + (p is None)

    def _concrete_table(self):
        if self._nparams:
            raise ValueError("No model parameters specified.")
        if self._table is None:
            self._table = B.ModelTable.binomial([self._n], [self._p])
        return self._table

    def _split(self, params):
        if self._nparams == 0:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != self._nparams:
            raise ValueError(f"Wrong number of model parameters: expected {self._nparams}, got {len(params)}.")
        it = iter(params)
        ns = ps = None
        if self._n is None:
            ns = np.asarray(next(it))
            if ns.ndim != 1 or ns.dtype != np.int32:
                raise TypeError("n must be a rank-1 numpy array with dtype int32")
        if self._p is None:
            ps = _float_param(next(it))
        size = ns.size if ns is not None else ps.size
        if ns is None:
            ns = np.full(size, self._n, dtype=np.int32)
        if ps is None:
            ps = np.full(size, self._p)
        if ns.size != ps.size:
            raise ValueError("Model parameters have unequal shape")
        return ns, ps

    def _family_len(self, params):
        return self._split(params)[0].size

    def _family_table(self, params):
        ns, ps = self._split(params)
        return B.ModelTable.binomial(ns, ps)


class CustomModel(Model):
    """pybindings/stream/model.rs:150-260 + model/internals.rs:255-420: a model defined by a Python `cdf(x, *params)`
    callback (`approximate_inverse_cdf` only seeds the reference's search and never changes a result, quantize.rs:580-779).
    The callback is the user's Python code either way; here it is evaluated once per table entry at the half-integers,
    pushed through the leaky quantiser's arithmetic (quantize.rs:525-568: trunc(free_weight * cdf(s - 0.5)) + slack) and
    the resulting CDF rows are uploaded (one row per model, or per symbol for a family)."""

    def __init__(self, cdf, approximate_inverse_cdf, min_symbol_inclusive, max_symbol_inclusive):
        self._cdf = cdf
        self._ppf = approximate_inverse_cdf
        self._lo, self._hi = int(min_symbol_inclusive), int(max_symbol_inclusive)
        if not self._hi > self._lo:
            raise ValueError("max_symbol_inclusive must be greater than min_symbol_inclusive")
        if self._hi - self._lo >= (1 << 24):
            raise ValueError("support too large for 24-bit probabilities")
        self._nparams = None  # any number of per-symbol parameter arrays

    def _row(self, args) -> np.ndarray:
        n = self._hi - self._lo + 1
        free_weight = float((1 << 24) - 1 - (self._hi - self._lo))
        row = np.empty(n + 1, dtype=np.uint32)
        row[0], row[n] = 0, 1 << 24
        for i in range(1, n):
            v = free_weight * float(self._cdf(float(self._lo + i) - 0.5, *args))
            q = 0 if not v > 0.0 else (0xFFFFFFFF if v >= 4294967295.0 else int(v))  # Rust `f64 as u32`
            row[i] = (q + i) & 0xFFFFFFFF
        return row

    def _concrete_table(self):
        if self._table is None:
            self._table = B.ModelTable.from_cdf(self._row(())[None, :], self._lo)
        return self._table

    def _columns(self, params):
        cols = [_float_param(p) for p in params]
        if any(c.size != cols[0].size for c in cols):
            raise ValueError("Model parameters have unequal lengths.")
        return cols

    def _family_len(self, params):
        return self._columns(params)[0].size

    def _family_table(self, params):
        cols = self._columns(params)
        rows = np.stack([self._row(tuple(float(c[i]) for c in cols)) for i in range(cols[0].size)]) if cols[0].size else \
            self._row(())[None, :]
        return B.ModelTable.from_cdf(rows, self._lo)


class ScipyModel(CustomModel):
    """pybindings/stream/model.rs:262-345: wraps a `scipy.stats` distribution (or frozen distribution)."""

    def __init__(self, scipy_model, min_symbol_inclusive, max_symbol_inclusive):
        super().__init__(scipy_model.cdf, scipy_model.ppf, min_symbol_inclusive, max_symbol_inclusive)
