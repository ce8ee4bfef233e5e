"""Shared by tests/golden/make_golden_modules.py (which runs the reference's modules on CPU) and
tests/test_modules_golden_gpu.py (which replays the same steps through chipmunk_b200 on the GPU): how the inputs of every
step are derived from the stored base tensors.  Integer hashing instead of an RNG, fp32 arithmetic on the CPU: both sides
build bit-identical bf16 inputs whatever the torch version."""
import numpy as np
import torch

BF = torch.bfloat16
ROW_STRIDE = 3          # outputs are stored for every third row (every 192-row group and 128-row block is sampled evenly)


def det_noise(shape, salt: int) -> torch.Tensor:
    """Uniform [-0.5, 0.5) values from an integer hash of (position, salt); exact in fp32."""
    n = int(np.prod(shape))
    idx = torch.arange(n, dtype=torch.int64)
    h = (idx * 2654435761 + (salt + 1) * 40503) % 2147483647
    h = (h * 48271 + 11) % 2147483647
    return ((h % 65536).to(torch.float32) / 65536.0 - 0.5).reshape(shape)


def to_bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().cpu().view(torch.int16).numpy().view(np.uint16)


def from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(BF)


def attn_step_inputs(q0, k0, v0, s: int, salt: int):
    """q, k, v of inference step s: the base tensors plus a drift that grows with the step."""
    d = 0.5 * s
    q = (q0.float() + d * det_noise(q0.shape, salt + 10 * s + 1)).to(BF)
    k = (k0.float() + 0.5 * d * det_noise(k0.shape, salt + 10 * s + 2)).to(BF)
    v = (v0.float() + 0.5 * d * det_noise(v0.shape, salt + 10 * s + 3)).to(BF)
    return q, k, v


def mlp_step_input(x0, dirs, s: int):
    """Tokens of step s: every 128-token block drifts along its own direction (fc1 is built so that only the block's
    chosen neurons see that direction), plus a little noise everywhere."""
    x = x0.float().clone()
    for b in range(x.shape[1] // 128):
        x[0, b * 128:(b + 1) * 128] += (0.8 * s) * dirs[b][None, :]
    x += 0.03 * s * det_noise(x.shape, 77 + s)
    return x.to(BF)
