"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's chunked forward + decode path.

Nothing in `bonito_b200/` may import this package; only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / reference arm do, and only as the checker / the thing timed as "CPU".

Parity status:
  * forward (conv -> LSTM stack -> linear -> clamp): PINNED against the reference's own modules
    (`/root/reference/bonito/nn.py`) imported in the authoring container -- `oracle/make_golden.py`
    writes `tests/golden/*.npz`, `tests/test_oracle_golden.py` replays them without the reference.
  * state graph / viterbi path logic: PINNED against `bonito/crf/model.py` (CTC_CRF.idx, viterbi).
  * posteriors / logZ: the reference delegates them to the closed-source `koi.ctc` (ont-koi==0.5.4,
    requirements.txt:19), absent here and unpinned by any reference test => "parity unpinned" for
    that step: the restatement follows the published definition used by the call sites
    (`bonito/crf/model.py:47-67`): logZ of the sparse transition graph in the Log / Max semiring and
    posteriors = d logZ / d scores, cross-checked against autograd in tests/test_oracle_crf.py.
  * koi.decode.beam_search (the CLI decoder): closed source, unpinned; not restated.  The decode oracle
    is the in-repo decode definition `SeqdistModel.decode_batch` (bonito/crf/model.py:196-199).
"""

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# forward path
# --------------------------------------------------------------------------------------------------

_ACT = {
    None: lambda x: x,
    "swish": F.silu,            # bonito/nn.py:54-56  Swish = SiLU
    "tanh": torch.tanh,
    "relu": torch.relu,
}


def _r16(x, fp16):
    """Round to fp16 and back: the storage rounding points of the reference's half-precision path."""
    return x.half().float() if fp16 else x


def convolution(x, weight, bias, stride, padding, activation, fp16=False):
    """Conv1d -> activation (BN already folded).  bonito/nn.py:235-241."""
    return _r16(_ACT[activation](_r16(F.conv1d(x, weight, bias, stride=stride, padding=padding), fp16)), fp16)


def lstm_layer(x, w_ih, w_hh, b_ih, b_hh, reverse, fp16=False):
    """
    Single-layer unidirectional LSTM over x [T, N, I], zero initial state, gate order i,f,g,o
    (torch.nn.LSTM semantics); `reverse` = flip, run, flip (bonito/nn.py:366-370).
    With `fp16` the input projection and every h_t are rounded to fp16 (fp32 accumulation and cell state),
    the rounding points of a half-precision LSTM.
    """
    T, N, _ = x.shape
    H = w_hh.shape[1]
    if reverse:
        x = x.flip(0)
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    gx = _r16(x @ w_ih.T + (b_ih + b_hh), fp16)
    out = []
    for t in range(T):
        g = gx[t] + h @ w_hh.T
        i, f, gg, o = g.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = _r16(torch.sigmoid(o) * torch.tanh(c), fp16)
        out.append(h)
    y = torch.stack(out)
    return y.flip(0) if reverse else y


def linear_crf(x, weight, bias, activation=None, scale=None, blank_score=None, n_base=4, expand_blanks=True, fp16=False):
    """LinearCRFEncoder.forward (bonito/nn.py:283-298) on x [T, N, H]; fp16: every op's result is stored as fp16."""
    s = _r16(F.linear(x, weight, bias), fp16)
    if activation is not None:
        s = _r16(_ACT[activation](s), fp16)
    if scale is not None:
        s = _r16(s * scale, fp16)
    if blank_score is not None and expand_blanks:
        T, N, C = s.shape
        s = F.pad(s.view(T, N, C // n_base, n_base), (1, 0, 0, 0, 0, 0, 0, 0), value=blank_score).view(T, N, -1)
    return s


def lstm_crf_forward(weights, spec, x, expand_blanks=False, return_features=False, fp16=False):
    """
    Whole LSTM-CRF encoder in fp32 on CPU.

    weights: dict produced by `oracle.synth.make_weights` (torch-layout fp32 tensors)
    spec:    dict(convs=[(cin,cout,k,stride,pad,act)...], hidden, n_lstm, reverse=[...], state_len, blank_score, clamp)
    x:       [N, 1, L] float32
    Returns scores [T, N, C] (reference layout; blanks expanded when `expand_blanks`).
    `fp16=True` keeps fp32 arithmetic but rounds every stored activation to fp16 (what `model.half()` stores).
    """
    feats = {}
    h = x
    for i, (_, _, _, stride, pad, act) in enumerate(spec["convs"]):
        h = convolution(h, weights[f"conv{i}.weight"], weights[f"conv{i}.bias"], stride, pad, act, fp16)
        feats[f"conv{i}"] = h
    h = h.permute(2, 0, 1)  # Permute([2,0,1]): NCT -> TNC
    for i in range(spec["n_lstm"]):
        h = lstm_layer(h, weights[f"lstm{i}.w_ih"], weights[f"lstm{i}.w_hh"], weights[f"lstm{i}.b_ih"],
                       weights[f"lstm{i}.b_hh"], spec["reverse"][i], fp16)
        feats[f"lstm{i}"] = h
    s = linear_crf(_r16(h, fp16), weights["crf.weight"], weights.get("crf.bias"), activation=spec.get("crf_activation"),
                   scale=spec.get("crf_scale"), blank_score=spec["blank_score"], expand_blanks=False, fp16=fp16)
    s = _r16(s, fp16)
    if expand_blanks:
        T_, N_, C_ = s.shape
        s = F.pad(s.view(T_, N_, C_ // 4, 4), (1, 0), value=spec["blank_score"]).view(T_, N_, -1)
    if spec.get("clamp") is not None:
        s = s.clamp(*spec["clamp"])  # Clamp: bonito/nn.py:66-67
    return (s, feats) if return_features else s


# --------------------------------------------------------------------------------------------------
# CRF state graph and semiring DP  (bonito/crf/model.py:30-67, 98-108)
# --------------------------------------------------------------------------------------------------

def crf_idx(state_len, n_base=4):
    """idx[s, 0] = s ; idx[s, 1 + j] = j * n_base**(state_len-1) + s // n_base  (bonito/crf/model.py:37-42)."""
    S = n_base ** state_len
    s = np.arange(S)
    idx = np.empty((S, n_base + 1), dtype=np.int64)
    idx[:, 0] = s
    for j in range(n_base):
        idx[:, 1 + j] = j * (S // n_base) + s // n_base
    return idx


def expand_blanks(scores, blank_score, n_base=4):
    """[..., S*n_base] -> [..., S*(n_base+1)] with the fixed stay score first in each group."""
    lead = scores.shape[:-1]
    x = scores.reshape(*lead, -1, n_base)
    pad = np.full((*lead, x.shape[-2], 1), blank_score, dtype=x.dtype)
    return np.concatenate([pad, x], axis=-1).reshape(*lead, -1)


def _lse(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def fwd_bwd(Ms, idx, semiring="log"):
    """
    Ms [T, N, S, E] edge scores (E = n_base+1 in-edges of each state), alpha_0 = beta_T = 0.
    Returns alpha [T+1, N, S], beta [T+1, N, S] with
        alpha_{t+1}[s] = (+)_e  Ms[t, s, e] (x) alpha_t[idx[s, e]]
        beta_t[p]      = (+)_{(s,e): idx[s,e]=p}  Ms[t, s, e] (x) beta_{t+1}[s]
    in the Log ((+) = logsumexp) or Max semiring.  Restates koi.ctc.fwd/bwd_scores_cu_sparse as used at
    bonito/crf/model.py:57-67.
    """
    T, N, S, E = Ms.shape
    red = _lse if semiring == "log" else (lambda x, axis: x.max(axis=axis))
    alpha = np.zeros((T + 1, N, S), dtype=Ms.dtype)
    beta = np.zeros((T + 1, N, S), dtype=Ms.dtype)
    for t in range(T):
        alpha[t + 1] = red(Ms[t] + alpha[t][:, idx], axis=-1)
    # successors of p: gather all (s, e) with idx[s, e] == p
    succ_s = [[] for _ in range(S)]
    succ_e = [[] for _ in range(S)]
    for s in range(S):
        for e in range(E):
            succ_s[idx[s, e]].append(s)
            succ_e[idx[s, e]].append(e)
    succ_s = np.array(succ_s)
    succ_e = np.array(succ_e)  # [S, E] (every state has exactly E successors in this graph)
    for t in range(T - 1, -1, -1):
        beta[t] = red(Ms[t][:, succ_s, succ_e] + beta[t + 1][:, succ_s], axis=-1)
    return alpha, beta


def logZ(Ms, idx, semiring="log"):
    alpha, _ = fwd_bwd(Ms, idx, semiring)
    red = _lse if semiring == "log" else (lambda x, axis: x.max(axis=axis))
    return red(alpha[-1], axis=-1)


def posteriors(Ms, idx):
    """Log-semiring edge marginals = d logZ / d Ms, shape [T, N, S, E]."""
    alpha, beta = fwd_bwd(Ms, idx, "log")
    lz = _lse(alpha[-1], axis=-1)  # [N]
    x = alpha[:-1][:, :, idx] + Ms + beta[1:][:, :, :, None] - lz[None, :, None, None]
    return np.exp(x)


def viterbi_edges(lp, idx):
    """
    Best path through edge scores lp [T, N, S, E] in the Max semiring: returns (state [T, N], edge [T, N]) =
    the destination state and in-edge taken at every frame.  Tie-breaks: lowest edge index, lowest final state.
    (The reference takes argmax over koi's Max-semiring gradient, bonito/crf/model.py:98-100.)
    """
    T, N, S, E = lp.shape
    v = np.zeros((N, S), dtype=lp.dtype)
    bp = np.empty((T, N, S), dtype=np.int64)
    for t in range(T):
        cand = lp[t] + v[:, idx]
        bp[t] = cand.argmax(axis=-1)  # first maximum = lowest edge
        v = np.take_along_axis(cand, bp[t][..., None], axis=-1)[..., 0]
    state = v.argmax(axis=-1)
    states = np.empty((T, N), dtype=np.int64)
    edges = np.empty((T, N), dtype=np.int64)
    rows = np.arange(N)
    for t in range(T - 1, -1, -1):
        e = bp[t, rows, state]
        states[t], edges[t] = state, e
        state = idx[state, e]
    return states, edges


def decode_batch(scores_tnc_blank, state_len, alphabet="NACGT", n_base=4):
    """
    SeqdistModel.decode_batch (bonito/crf/model.py:196-199) on blank-expanded scores [T, N, S*5]:
        post = posteriors(scores.float()) + 1e-8 ; path = viterbi(log post) ; strings.
    Returns (strings, path [N, T] with 0 = no emission else 1 + base).
    """
    T, N, C = scores_tnc_blank.shape
    idx = crf_idx(state_len, n_base)
    Ms = scores_tnc_blank.astype(np.float64).reshape(T, N, -1, n_base + 1)
    post = posteriors(Ms, idx).astype(np.float32)
    lp = np.log(post + np.float32(1e-8))
    states, edges = viterbi_edges(lp, idx)
    # viterbi(): a = s*5 + e ; move = e != 0 ; base = 1 + (a // 5) % 4   (bonito/crf/model.py:101-103)
    path = np.where(edges != 0, 1 + states % n_base, 0).T  # [N, T]
    letters = np.frombuffer(alphabet.encode(), dtype="u1")
    strings = [letters[p[p != 0]].tobytes().decode() for p in path]
    return strings, path


def decode_native(scores_ntc, state_len, blank_score=2.0, qscale=1.0, qbias=0.0, n_base=4):
    """
    Oracle for b200_crf_decode: scores [N, T, S*4] (no blanks, fp16-valued) ->
    (moves, sequence, qstring) uint8 [N, T] with the kernel's output conventions, and the move-mass table.
    """
    x = np.asarray(scores_ntc, dtype=np.float64)
    N, T, C = x.shape
    idx = crf_idx(state_len, n_base)
    Ms = expand_blanks(x.transpose(1, 0, 2), blank_score, n_base).reshape(T, N, -1, n_base + 1)
    post = posteriors(Ms, idx)
    lp = np.log(post.astype(np.float32) + np.float32(1e-8))
    states, edges = viterbi_edges(lp, idx)
    move = edges != 0
    base = states % n_base
    # posterior mass of "a move that emits base b" per frame
    S = Ms.shape[2]
    mass = post[..., 1:].sum(-1).reshape(T, N, S // n_base, n_base).sum(2)  # [T, N, 4]
    p = np.take_along_axis(mass, base[..., None], axis=-1)[..., 0]
    err = np.maximum(1.0 - p, 1e-4)
    q = np.rint(-10.0 * np.log10(err) * qscale + qbias).astype(np.int64) + 33  # util.phred, bonito/util.py:105-112
    q = np.clip(q, 33, 126)
    letters = np.frombuffer(b"ACGT", dtype="u1")
    moves = move.T.astype(np.uint8)
    seq = np.where(move, letters[base], 0).T.astype(np.uint8)
    qual = np.where(move, q, 0).T.astype(np.uint8)
    return moves, seq, qual, mass.transpose(1, 0, 2)


# --------------------------------------------------------------------------------------------------
# beam search  (koi.decode.beam_search call contract, bonito/crf/basecall.py:36-40)
# --------------------------------------------------------------------------------------------------
# koi's decoder is a closed binary with no pinned outputs ("parity unpinned"), so this restates
# the ALGORITHM THIS REPOSITORY SHIPS behind the same signature -- a backward-guided prefix beam search over the k-mer
# CRF -- and is the checker of the CUDA kernel, not a parity claim against koi:
#   * a beam entry is a (sequence, k-mer state) pair: identified by a 64-bit hash of the emitted bases + start state;
#   * per frame every entry spawns "stay" (blank score) and four "move" candidates (state' = (state % Q) * 4 + base, score
#     M_t[state' * 4 + state / Q]); a stay candidate and a move candidate that reach the same sequence are merged by
#     log-add (the two alignments of that prefix), recording the larger one as the back-pointer;
#   * candidates are ranked by forward score + beta'_{t+1}[state'] (the exact backward scores of the forward-backward pass
#     as look-ahead), candidates more than `beam_cut` below the best are dropped, the best `beam_width` survive
#     (ties: lower candidate index = lower parent entry, stay before moves, lower base);
#   * the start beam is the `beam_width` best start states by beta'_0; the answer is the trace-back of the best final entry.
BEAM_HASH_MULT = np.uint64(0x9E3779B97F4A7C15)


def beam_search_native(scores_ntc, state_len, blank_score=2.0, beam_width=32, beam_cut=100.0, n_base=4):
    """scores [N, T, S*4] (no blanks) -> (moves, bases) uint8 [N, T]; bases: 0 = no emission, else 1 + base."""
    x = np.asarray(scores_ntc, dtype=np.float32)
    N, T, C = x.shape
    S = n_base ** state_len
    Q = S // n_base
    idx = crf_idx(state_len, n_base)
    Ms = expand_blanks(x.astype(np.float64).transpose(1, 0, 2), blank_score, n_base).reshape(T, N, S, n_base + 1)
    _, beta = fwd_bwd(Ms, idx, "log")                         # [T+1, N, S]
    beta = (beta - beta[:, :, :1]).astype(np.float32)         # re-centred on state 0 per frame, as the kernel keeps it
    blank = np.float32(blank_score)
    cut = np.float32(beam_cut)
    moves = np.zeros((N, T), dtype=np.uint8)
    bases = np.zeros((N, T), dtype=np.uint8)
    W = beam_width
    for n in range(N):
        order = np.argsort(-beta[0, n], kind="stable")[:W]
        bh = (order.astype(np.uint64) + np.uint64(1))
        bs = order.astype(np.int64)
        bsc = np.zeros(len(order), dtype=np.float32)
        bp_parent = np.zeros((T, W), dtype=np.int64)
        bp_code = np.zeros((T, W), dtype=np.int64)
        counts = np.zeros(T, dtype=np.int64)
        for t in range(T):
            nb = len(bs)
            m = x[n, t]
            # candidates in index order: entry-major, slot 0 = stay, 1 + b = move with base b
            ch = np.empty((nb, 5), dtype=np.uint64)
            cs = np.empty((nb, 5), dtype=np.int64)
            csc = np.empty((nb, 5), dtype=np.float32)
            ch[:, 0], cs[:, 0], csc[:, 0] = bh, bs, bsc + blank
            with np.errstate(over="ignore"):
                for b in range(4):
                    s2 = (bs % Q) * 4 + b
                    ch[:, 1 + b] = bh * BEAM_HASH_MULT + np.uint64(b + 1)
                    cs[:, 1 + b] = s2
                    csc[:, 1 + b] = bsc + m[s2 * 4 + bs // Q]
            alive = np.ones((nb, 5), dtype=bool)
            code = np.tile(np.arange(5), (nb, 1))
            parent = np.tile(np.arange(nb)[:, None], (1, 5))
            # merge: the stay candidate of entry j with the move candidate that spells the same sequence
            for j in range(nb):
                hit = np.argwhere((ch[:, 1:] == ch[j, 0]) & alive[:, 1:])
                if len(hit):
                    i, b = hit[0]
                    a, c = csc[j, 0], csc[i, 1 + b]
                    hi_, lo_ = (a, c) if a >= c else (c, a)
                    csc[j, 0] = np.float32(hi_ + np.log1p(np.exp(np.float32(lo_ - hi_), dtype=np.float32), dtype=np.float32))
                    if c > a:
                        code[j, 0], parent[j, 0] = 1 + b, i
                    alive[i, 1 + b] = False
            key = np.where(alive, csc + beta[t + 1, n][cs], -np.inf).astype(np.float32)
            best = key.max()
            key = np.where(key < best - cut, -np.inf, key)
            flat = np.argsort(-key.reshape(-1), kind="stable")
            flat = flat[np.isfinite(key.reshape(-1)[flat])][:W]
            sel_i, sel_c = flat // 5, flat % 5
            new_sc = csc[sel_i, sel_c]
            counts[t] = len(flat)
            bp_parent[t, :len(flat)] = parent[sel_i, sel_c]
            bp_code[t, :len(flat)] = code[sel_i, sel_c]
            bh, bs = ch[sel_i, sel_c], cs[sel_i, sel_c]
            bsc = (new_sc - new_sc.max()).astype(np.float32)
        r = int(np.argmax(bsc))           # first maximum = lowest entry
        for t in range(T - 1, -1, -1):
            c = bp_code[t, r]
            if c:
                moves[n, t] = 1
                bases[n, t] = c           # 1 + base
            r = bp_parent[t, r]
    return moves, bases
