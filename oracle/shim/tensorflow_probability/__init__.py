"""TEST INFRASTRUCTURE ONLY (oracle shim): the one TFP-0.18 distribution the reference uses
(MixtureGPs/models.py:60,73,94) plus the two bijectors GPflow parameters use.  [3P-memory]:
tensorflow_probability/python/distributions/relaxed_onehot_categorical.py
(ExpRelaxedOneHotCategorical._sample_n followed by the Exp bijector), SURVEY.md Appendix A.6/A.7."""
import numpy as _np
import torch as _torch
import tensorflow as _tf


class _RelaxedOneHotCategorical:
    def __init__(self, temperature, logits=None, probs=None):
        assert logits is not None
        self.temperature = _tf._t(temperature)
        self.logits = _tf._t(logits)

    def sample(self, n, seed=None):
        n = int(n) if not isinstance(n, (list, tuple)) else int(n[0])
        shp = [n] + list(self.logits.shape)
        # uniform on (tiny, 1): open at 0 so that log(-log u) is finite
        uniform = _tf.random.uniform(shp, minval=_np.finfo(_np.float64).tiny, maxval=1.0)
        gumbel = -_torch.log(-_torch.log(uniform))
        noisy_logits = (gumbel + self.logits) / self.temperature[..., None]
        return _torch.exp(_torch.log_softmax(noisy_logits, dim=-1))


class distributions:  # noqa: N801
    RelaxedOneHotCategorical = _RelaxedOneHotCategorical


class _Softplus:
    def forward(self, x):
        return _torch.nn.functional.softplus(x)

    def inverse(self, y):
        # tfp Softplus._inverse: log(expm1(y)) written stably as y + log(-expm1(-y))
        y = _torch.as_tensor(y, dtype=_torch.float64)
        return y + _torch.log(-_torch.expm1(-y))


class _Sigmoid:
    def forward(self, x):
        return _torch.sigmoid(x)

    def inverse(self, y):
        y = _torch.as_tensor(y, dtype=_torch.float64)
        return _torch.log(y) - _torch.log1p(-y)


def _fill_triangular_index(m):
    """TFP fill_triangular (lower): x (len m(m+1)/2) -> concat([x[m:], reversed(x)]) reshaped [m,m], tril.
    Returns flat gather indices idx so that tril_flat = x[idx] on the lower triangle (-1 elsewhere)."""
    n = m * (m + 1) // 2
    x = _np.arange(n)
    xc = _np.concatenate([x[m:], x[::-1]])
    mat = xc.reshape(m, m)
    idx = _np.where(_np.tril(_np.ones((m, m), dtype=bool)), mat, -1)
    return idx


class _FillTriangular:
    def forward(self, x):
        x = _torch.as_tensor(x, dtype=_torch.float64)
        n = x.shape[-1]
        m = int((_np.sqrt(8 * n + 1) - 1) / 2)
        idx = _fill_triangular_index(m)
        mask = _torch.as_tensor(idx >= 0)
        gathered = x[..., _torch.as_tensor(_np.maximum(idx, 0)).reshape(-1)].reshape(*x.shape[:-1], m, m)
        return _torch.where(mask, gathered, _torch.zeros((), dtype=x.dtype))

    def inverse(self, y):
        y = _torch.as_tensor(y, dtype=_torch.float64)
        m = y.shape[-1]
        idx = _fill_triangular_index(m)
        n = m * (m + 1) // 2
        out = _torch.zeros(*y.shape[:-2], n, dtype=y.dtype)
        ii, jj = _np.nonzero(idx >= 0)
        out[..., _torch.as_tensor(idx[ii, jj])] = y[..., _torch.as_tensor(ii), _torch.as_tensor(jj)]
        return out


class bijectors:  # noqa: N801
    Softplus = _Softplus
    Sigmoid = _Sigmoid
    FillTriangular = _FillTriangular
