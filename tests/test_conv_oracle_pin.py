"""Second, independent pin of oracle/conv_oracle.py (not a GPU test): the flipped-kernel torch conv2d it uses is checked against a direct
loop transcription of the NNlib.conv definition Flux's Conv calls [3P] -- y[ow, oh, co, b] = sum_{kw, kh, ci} w[kw, kh, ci, co] *
x[s*(ow-1) + (KW - kw + 1), s*(oh-1) + (KH - kh + 1), ci, b] (1-based, no padding) -- computed on column-major arrays exactly as Julia lays
them out, and Flux.flatten against a column-major reshape."""
import numpy as np
import torch

from oracle import conv_oracle as co

F32 = np.float32


def _nnlib_conv_loop(x_whcn, w_kkio, b, stride):
    """x [W, H, C, B], w [KW, KH, Cin, Cout] (Julia axes), true convolution, no padding."""
    W, H, C, B = x_whcn.shape
    KW, KH, _, CO = w_kkio.shape
    OW, OH = (W - KW) // stride + 1, (H - KH) // stride + 1
    y = np.zeros((OW, OH, CO, B), dtype=np.float64)
    for ow in range(1, OW + 1):
        for oh in range(1, OH + 1):
            for kw in range(1, KW + 1):
                for kh in range(1, KH + 1):
                    xi = x_whcn[stride * (ow - 1) + (KW - kw + 1) - 1, stride * (oh - 1) + (KH - kh + 1) - 1]    # [C, B]
                    y[ow - 1, oh - 1] += np.einsum("cb,co->ob", xi.astype(np.float64), w_kkio[kw - 1, kh - 1].astype(np.float64))
    return y + b.reshape(1, 1, -1, 1)


def test_conv_oracle_equals_the_nnlib_definition():
    rng = np.random.default_rng(0)
    net = co.ConvQ(3, 11, 9, 3, 2, 4, 2, 1, 5, 6, 2, rng)          # C, H, W, k1, s1, c1, k2, s2, c2, hidden, nA
    s = rng.integers(0, 256, (2, 3, 11, 9), dtype=np.uint8)         # memory [b][c][h][w]
    # the same memory as Julia arrays: x[w, h, c, b], W[kw, kh, ci, co]
    x = (s.astype(F32) / F32(255)).transpose(3, 2, 1, 0)
    W1 = net.W1.detach().numpy().transpose(3, 2, 1, 0)
    W2 = net.W2.detach().numpy().transpose(3, 2, 1, 0)
    y1 = np.maximum(_nnlib_conv_loop(x, W1, net.b1.detach().numpy(), 2), 0)
    y2 = np.maximum(_nnlib_conv_loop(y1, W2, net.b2.detach().numpy(), 1), 0)
    flat = y2.reshape(-1, y2.shape[-1], order="F")                  # Flux.flatten: reshape(x, :, B) of the column-major array
    h = np.maximum(net.head.W[0].detach().numpy().astype(np.float64) @ flat + net.head.b[0].detach().numpy().reshape(-1, 1), 0)
    q = net.head.W[1].detach().numpy().astype(np.float64) @ h + net.head.b[1].detach().numpy().reshape(-1, 1)
    np.testing.assert_allclose(net(s).detach().numpy(), q.T, rtol=2e-5, atol=2e-6)


def test_clip_value_then_adam():
    """Flux.Optimiser(ClipValue(1), Adam): the first Adam step moves every coordinate by eta * sign(clamped g) (bias-corrected m / sqrt(v) = +-1),
    and a coordinate whose raw gradient exceeds the threshold is treated as having gradient +-1 in the moments."""
    p = torch.tensor([0.0, 0.0, 0.0], requires_grad=True)
    p.grad = torch.tensor([5.0, -0.5, 0.0])
    opt = co.ClipAdam(1.0, F32(1e-3))
    opt.step([p])
    m, v, _ = opt.state[id(p)]
    np.testing.assert_allclose(m, [0.1, -0.05, 0.0], rtol=1e-6)
    np.testing.assert_allclose(v, [1e-3, 0.25e-3, 0.0], rtol=1e-5)
    np.testing.assert_allclose(p.detach().numpy(), [-1e-3, 1e-3, 0.0], rtol=1e-4, atol=1e-12)
