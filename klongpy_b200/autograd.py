"""Autograd interop — placeholder until the differentiable wrappers land.

The reference's torch backend differentiates by running the interpreter on `requires_grad` tensors
(`torch_backend.py:692-782`).  The B200 path will wrap the fused kernels in `torch.autograd.Function`s; until
then the hooks raise, so `:>` reports a clear error instead of silently computing on the CPU.
"""


def _todo(*_a, **_k):
    raise NotImplementedError("klongpy_b200: autograd interop is not implemented yet (SURVEY.md §8 row a14)")


create_grad_tensor = compute_autograd = compute_multi_autograd = compute_jacobian = _todo
