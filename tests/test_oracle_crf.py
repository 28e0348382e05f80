"""Self-consistency of the CRF restatement: the closed-source koi.ctc pieces are pinned by definition
(logZ of the sparse graph; posteriors = d logZ / d scores) rather than by reference outputs."""
import itertools

import numpy as np
import torch

from oracle import crf_oracle as O


def _brute_force_logz(Ms, idx):
    """Enumerate every state path of a tiny problem."""
    T, N, S, E = Ms.shape
    out = np.zeros(N)
    for n in range(N):
        tot = []
        for path in itertools.product(range(S), repeat=T + 1):
            sc = 0.0
            ok = True
            for t in range(T):
                prev, cur = path[t], path[t + 1]
                edges = [e for e in range(E) if idx[cur, e] == prev]
                if not edges:
                    ok = False
                    break
                # parallel edges (e.g. AAAA -> AAAA by stay or by move) both count
                sc_t = np.logaddexp.reduce([Ms[t, n, cur, e] for e in edges])
                sc += sc_t
            if ok:
                tot.append(sc)
        out[n] = np.logaddexp.reduce(tot)
    return out


def test_logz_against_path_enumeration():
    rng = np.random.default_rng(0)
    idx = O.crf_idx(1)  # 4 states
    Ms = rng.normal(size=(4, 2, 4, 5))
    np.testing.assert_allclose(O.logZ(Ms, idx), _brute_force_logz(Ms, idx), rtol=1e-10)


def test_posteriors_are_the_gradient_of_logz():
    rng = np.random.default_rng(1)
    k = 2
    idx = O.crf_idx(k)
    Ms = rng.normal(size=(6, 3, 16, 5))
    post = O.posteriors(Ms, idx)
    x = torch.tensor(Ms, requires_grad=True)
    tidx = torch.from_numpy(idx)
    alpha = torch.zeros(3, 16, dtype=torch.float64)
    for t in range(6):
        alpha = torch.logsumexp(x[t] + alpha[:, tidx], dim=-1)
    torch.logsumexp(alpha, dim=-1).sum().backward()
    np.testing.assert_allclose(post, x.grad.numpy(), atol=1e-12)
    np.testing.assert_allclose(post.sum(axis=(2, 3)), 1.0, atol=1e-12)  # one transition per frame


def test_alpha_beta_consistency_and_max_semiring():
    rng = np.random.default_rng(2)
    idx = O.crf_idx(3)
    Ms = rng.normal(size=(20, 2, 64, 5)) * 2
    for semi in ("log", "max"):
        a, b = O.fwd_bwd(Ms, idx, semi)
        red = O._lse if semi == "log" else (lambda v, axis: v.max(axis=axis))
        z = red(a[-1], axis=-1)
        for t in (0, 7, 20):
            np.testing.assert_allclose(red(a[t] + b[t], axis=-1), z, rtol=1e-12)
    # Viterbi path score equals the Max-semiring logZ
    states, edges = O.viterbi_edges(Ms, idx)
    t_idx, n_idx = np.meshgrid(np.arange(20), np.arange(2), indexing="ij")
    np.testing.assert_allclose(Ms[t_idx, n_idx, states, edges].sum(0), O.logZ(Ms, idx, "max"), rtol=1e-12)
    # path is connected: predecessor of state[t] along edge[t] is state[t-1]
    assert np.array_equal(idx[states[1:], edges[1:]], states[:-1])


def test_expand_blanks_layout():
    x = np.arange(2 * 3 * 16, dtype=np.float32).reshape(2, 3, 16)
    y = O.expand_blanks(x, 2.0).reshape(2, 3, 4, 5)
    assert np.all(y[..., 0] == 2.0) and np.array_equal(y[..., 1:].reshape(2, 3, 16), x)
