"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never import this from the product package.

Emulation, in torch float64 on the CPU, of the pieces of the un-vendored third-party library
``nisaba`` (gitlab.com/sci-learning/nisaba, no version pinned: README.md:27) and of
TensorFlow/Keras 2.7 that the reference's hot path executes in.  nisaba's source is not under
/root/reference, so its semantics are restated from the reference's own call sites and saved
artefacts; each item says where the evidence is.

PARITY UNPINNED for numerical loss/gradient values: TensorFlow and nisaba cannot be imported in
this image and the reference has no tests or golden vectors.  What IS pinned (tests/test_oracle_pins.py):
  * total loss == sum_t weight_t * value_t            (all History_Loss.json files, to 1e-15)
  * an out-of-tape ``divergence_vector`` logs exactly 0 (Colliding #003 / Poiseuille #016 histories)
  * the trained Weights.h5 of Colliding #003 / Poiseuille #016 evaluated by this oracle on the
    scripts' grids reproduce the recorded final loss values to within sampling scatter
  * analytic solutions give zero residuals through these operators
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

DTYPE = torch.float64  # ns.config.get_dtype(): Model.json "dtype": "float64"


class KerasMLP:
    """tf.keras.Sequential([Dense(H, tanh)] * L + [Dense(O)]) -- cavity_steady.py:205-210.

    ``variables`` follow Keras order [K1, b1, ..., K_{L+1}, b_{L+1}], kernels [in, out], y = x @ K + b.
    """

    def __init__(self, variables: Sequence[torch.Tensor]):
        self.variables = [v.detach().clone().to(DTYPE).requires_grad_(True) for v in variables]

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        h = x.to(DTYPE)
        n_layers = len(self.variables) // 2
        for l in range(n_layers):
            K, b = self.variables[2 * l], self.variables[2 * l + 1]
            h = h @ K + b
            if l < n_layers - 1:
                h = torch.tanh(h)
        return h


class GradientTape:
    """ns.GradientTape(persistent=True): thin wrapper of tf.GradientTape (Report.pdf p.27, B.1).

    ``active`` tracks whether we are inside the ``with`` block: TensorFlow records only the ops
    executed while the tape is open, which is what makes an operator called after the block see an
    unconnected graph (quirk Q1).
    """

    def __init__(self, persistent: bool = True):
        self.persistent = persistent
        self.active = False

    def __enter__(self):
        self.active = True
        return self

    def __exit__(self, *exc):
        self.active = False
        return False

    def watch(self, x: torch.Tensor) -> None:
        x.requires_grad_(True)

    def gradient(self, y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        # tf.GradientTape.gradient(y, x) = d sum(y) / d x; unconnected -> zeros [inferred from the
        # constant BCN_v_OUT2 log of Coronary #123 and the 0.0 PDE_MASS logs].
        if not y.requires_grad:
            return torch.zeros_like(x)
        (g,) = torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True,
                                   allow_unused=True)
        return torch.zeros_like(x) if g is None else g


# nisaba.experimental.physics.tens_style ---------------------------------------------------------

def gradient_scalar(tape: GradientTape, s: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """[N] or [N,1] scalar field -> [N, d] (call sites: cavity_steady.py:157,164-165,177-181)."""
    return tape.gradient(s, x)


def divergence_vector(tape: GradientTape, v: torch.Tensor, x: torch.Tensor, dim: int) -> torch.Tensor:
    """sum_i d v_i / d x_i (call sites: colliding_flow.py:165, poiseuille_flow.py:178).

    The column slices ``v[:, i]`` are taken INSIDE this function.  When the caller has already left
    the ``with`` block they are not recorded on the tape and the result is identically zero -- the
    recorded PDE_MASS log of Colliding_Flow #003 and Poiseuille_Flow #016 is 0.0 at all 1011
    entries, while the in-tape formulation of Examples_Old/Poiseuille/poiseuille.py:98-103 logs
    non-zero values.
    """
    n = x.shape[0]
    if not tape.active:
        return torch.zeros(n, dtype=x.dtype)
    out = torch.zeros(n, dtype=x.dtype)
    for i in range(dim):
        out = out + tape.gradient(v[:, i], x)[:, i]
    return out


def laplacian_scalar(tape: GradientTape, s: torch.Tensor, x: torch.Tensor, dim: int) -> torch.Tensor:
    """sum_i d2 s / d x_i^2 as divergence of the gradient (poisson.py:62, colliding_flow.py:179)."""
    if not tape.active:
        return torch.zeros(x.shape[0], dtype=x.dtype)
    g = gradient_scalar(tape, s, x)
    out = torch.zeros(x.shape[0], dtype=x.dtype)
    for i in range(dim):
        out = out + tape.gradient(g[:, i], x)[:, i]
    return out


# nisaba loss objects -------------------------------------------------------------------------------

class LossMeanSquares:
    """ns.LossMeanSquares(name, eval_roots, weight=1.0, normalization=1.0).

    value = mean(roots**2) / normalization.  Evidence: History_Loss.json stores per-term ``log``
    values whose weighted sum reproduces ``loss_global`` to 5e-16; ``normalization``
    appears only in colliding_flow_pressmean.py:184-186.
    """

    def __init__(self, name: str, eval_roots: Callable[[], torch.Tensor], weight: float = 1.0,
                 normalization: float = 1.0):
        self.name, self.eval_roots = name, eval_roots
        self.weight, self.normalization = float(weight), float(normalization)
        self.non_negative, self.display_sqrt = True, True

    def roots(self) -> torch.Tensor:
        return self.eval_roots()

    def __call__(self) -> torch.Tensor:
        r = self.eval_roots()
        return torch.mean(torch.square(r)) / self.normalization


class Loss:
    """ns.Loss(name, eval_loss, normalization=1.0, weight=1.0, non_negative=False): a scalar loss,
    value = eval_loss() / normalization [inferred by analogy with LossMeanSquares; the only call site,
    colliding_flow_pressmean.py:196, passes normalization=1e0]."""

    def __init__(self, name: str, eval_loss: Callable[[], torch.Tensor], normalization: float = 1.0, weight: float = 1.0,
                 non_negative: bool = False):
        self.name, self.eval_loss = name, eval_loss
        self.weight, self.normalization = float(weight), float(normalization)
        self.non_negative, self.display_sqrt = bool(non_negative), False

    def __call__(self) -> torch.Tensor:
        return self.eval_loss() / self.normalization


class OptimizationProblem:
    """ns.OptimizationProblem(variables, losses, losses_test): total = sum_t weight_t * loss_t()."""

    def __init__(self, variables: List[torch.Tensor], losses: Sequence[LossMeanSquares],
                 losses_test: Optional[Sequence[LossMeanSquares]] = None):
        self.variables = list(variables)
        self.losses = list(losses)
        if losses_test is None:
            losses_test = []
        elif isinstance(losses_test, LossMeanSquares):
            losses_test = [losses_test]  # poisson.py:69,72 passes a single loss object
        self.losses_test = list(losses_test)

    def loss_and_grad(self):
        """One evaluation of nisaba's step: values per term, total, d total / d variables."""
        values = [L() for L in self.losses]
        total = sum(L.weight * v for L, v in zip(self.losses, values))
        if total.requires_grad:
            grads = torch.autograd.grad(total, self.variables, allow_unused=True)
            grads = [torch.zeros_like(v) if g is None else g for g, v in zip(grads, self.variables)]
        else:
            grads = [torch.zeros_like(v) for v in self.variables]
        return ([float(v.detach()) for v in values], float(total.detach()),
                torch.cat([g.reshape(-1) for g in grads]).detach())

    def test_values(self):
        with torch.enable_grad():
            return [float(L().detach()) for L in self.losses_test]
