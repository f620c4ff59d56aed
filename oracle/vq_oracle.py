"""CPU restatement (numpy) of the semi-tts VQ bottleneck -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  Parity: PINNED by tests/golden/*.npz, which were produced
by the unmodified reference (oracle/gen_golden.py); see oracle/__init__.py.

Every function cites the reference lines it restates.  All functions take a `dtype`
(np.float32 to mirror the reference's arithmetic and evaluation order, np.float64 for
the high-precision yardstick that both the reference and the CUDA path are measured
against).

Notation: N = B*S rows (frames), D = latent_dim, K = vocab_size (codebook size),
A = number of phoneme attributes (31), D_a = proj_attr (16), D_l = D - D_a.
"""
import numpy as np


# --------------------------------------------------------------------------------------
# codebook table assembly                                     (src/embed.py:109-112, 87-94)
# --------------------------------------------------------------------------------------
def assemble_table(learnable_table, phn_attr=None, proj_w=None, proj_b=None, dtype=np.float64):
    """E[K,D] = cat([learnable_table[K,D_l], phn_attr[K,A] @ proj_w[D_a,A].T + proj_b], -1)."""
    lt = np.asarray(learnable_table, dtype=dtype)
    if phn_attr is None:
        return lt
    proj = np.asarray(phn_attr, dtype=dtype) @ np.asarray(proj_w, dtype=dtype).T \
        + np.asarray(proj_b, dtype=dtype)
    return np.concatenate([lt, proj], axis=-1)


# --------------------------------------------------------------------------------------
# distance                                                        (src/embed.py:208-213)
# --------------------------------------------------------------------------------------
def l2_distance(x2d, table, dtype=np.float64):
    """d[N,K] = (sum x^2 [N,1] + sum e^2 [K]) - 2 x e^T, evaluated in that order."""
    x = np.asarray(x2d, dtype=dtype)
    e = np.asarray(table, dtype=dtype)
    xx = np.sum(x * x, axis=-1, keepdims=True)
    ee = np.sum(e * e, axis=-1)
    return (xx + ee) - dtype(2) * (x @ e.T)


def softmax_rows(sim):
    m = np.max(sim, axis=-1, keepdims=True)
    z = np.exp(sim - m)
    return z / np.sum(z, axis=-1, keepdims=True)


def first_argmax(p):
    """torch.argmax on CPU returns the lowest index among tied maxima; so does np.argmax."""
    return np.argmax(p, axis=-1).astype(np.int64)


# --------------------------------------------------------------------------------------
# L2 quantizer forward                                            (src/embed.py:105-147)
# --------------------------------------------------------------------------------------
def l2_forward(x, table, temp, stop_grad=True, skip=False, dtype=np.float64):
    """x[B,S,D], table[K,D], temp scalar -> dict(p_code[B,S,K], idx[B,S], code[B,S,D], new_latent[B,S,D]).

    sim = relu(temp) * -d (:115-124; the real/fake split only changes gradients, not values),
    p_code = softmax(sim) (:127), idx = argmax(p_code) (:130), code = E[idx] (:134) or
    p_hard @ E (:137-138), new_latent = (x + code) - x (:145) or x when skipped (:142).
    """
    x = np.asarray(x, dtype=dtype)
    e = np.asarray(table, dtype=dtype)
    B, S, D = x.shape
    tau = dtype(max(float(temp), 0.0))
    d = l2_distance(x.reshape(B * S, D), e, dtype)
    sim = tau * (-d)
    p = softmax_rows(sim)
    idx = first_argmax(p)
    if stop_grad:
        code = e[idx]
    else:
        onehot = np.zeros_like(p)
        onehot[np.arange(B * S), idx] = 1
        p_hard = p + (onehot - p)
        code = p_hard @ e
    xf = x.reshape(B * S, D)
    new_latent = xf.copy() if skip else (xf + code) - xf
    K = e.shape[0]
    return dict(p_code=p.reshape(B, S, K), idx=idx.reshape(B, S), code=code.reshape(B, S, D),
                new_latent=new_latent.reshape(B, S, D), dist=d.reshape(B, S, K))


# --------------------------------------------------------------------------------------
# L2 quantizer backward (autograd of :105-147; algebra in SURVEY.md section 3.4)
# --------------------------------------------------------------------------------------
def l2_backward(x, table, temp, p_code, idx, g_p=None, g_q=None, stop_grad=True,
                first_n_real_rows=0, skip=False, dtype=np.float64):
    """Gradients of sum(p_code*g_p) + sum(new_latent*g_q).

    first_n_real_rows: number of leading flattened rows (first_n_real_mel * S) whose distance
    route reaches the table; 0 means ALL rows do (src/embed.py:115-124).
    Returns dict(dx[B,S,D], dtable[K,D], dtemp scalar).
    """
    x = np.asarray(x, dtype=dtype)
    e = np.asarray(table, dtype=dtype)
    B, S, D = x.shape
    N, K = B * S, e.shape[0]
    xf = x.reshape(N, D)
    P = np.asarray(p_code, dtype=dtype).reshape(N, K)
    ix = np.asarray(idx).reshape(N)
    tau = dtype(max(float(temp), 0.0))
    gp = np.zeros((N, K), dtype) if g_p is None else np.asarray(g_p, dtype=dtype).reshape(N, K)
    gq = np.zeros((N, D), dtype) if g_q is None else np.asarray(g_q, dtype=dtype).reshape(N, D)

    # straight-through identity term d new_latent / d x = I (also in the skip branch)
    dx = gq.copy()
    dtable = np.zeros((K, D), dtype)

    G = gp.copy()
    if not skip:
        if stop_grad:
            np.add.at(dtable, ix, gq)                 # F.embedding backward (:134)
        else:
            # p_hard = p + (onehot - p).detach(); code = p_hard @ E  (:137-138)
            G = G + gq @ e.T                          # d/dp_hard flows to p_code
            onehot = np.zeros((N, K), dtype)
            onehot[np.arange(N), ix] = 1
            dtable += onehot.T @ gq                   # value of p_hard is the one-hot
    s = np.sum(G * P, axis=-1, keepdims=True)
    Gs = P * (G - s)                                  # softmax backward (:127)
    Gd = -tau * Gs                                    # d L / d dist
    dx += dtype(2) * xf * np.sum(Gd, axis=-1, keepdims=True) - dtype(2) * (Gd @ e)
    Gd_tab = Gd
    if first_n_real_rows > 0:
        Gd_tab = Gd.copy()
        Gd_tab[first_n_real_rows:] = 0                # table was detached for those rows
    dtable += dtype(2) * e * np.sum(Gd_tab, axis=0)[:, None] - dtype(2) * (Gd_tab.T @ xf)
    d = l2_distance(xf, e, dtype)
    dtemp = np.sum(Gs * (-d)) if float(temp) > 0 else dtype(0)
    return dict(dx=dx.reshape(B, S, D), dtable=dtable, dtemp=dtemp)


def table_backward(dtable, phn_attr=None, proj_w=None, dtype=np.float64):
    """Split dE into parameter grads (autograd of :109-112).

    Returns dict(d_learnable[K,D_l], d_proj_w[D_a,A] or None, d_proj_b[D_a] or None)."""
    dt = np.asarray(dtable, dtype=dtype)
    if phn_attr is None:
        return dict(d_learnable=dt, d_proj_w=None, d_proj_b=None)
    Da = np.asarray(proj_w).shape[0]
    Dl = dt.shape[1] - Da
    g = dt[:, Dl:]
    return dict(d_learnable=dt[:, :Dl],
                d_proj_w=g.T @ np.asarray(phn_attr, dtype=dtype),
                d_proj_b=np.sum(g, axis=0))


# --------------------------------------------------------------------------------------
# inference (gather only)                                  (src/embed.py:96-103, 180-185)
# --------------------------------------------------------------------------------------
def inference(txt, table):
    """txt[B,L] int -> table[txt] [B,L,D]; identical for both quantizer variants because
    cat(embedding(txt), proj_attr(phn_attr(txt))) == cat(embedding.weight, proj_attr(phn_attr.weight))[txt]."""
    return np.asarray(table)[np.asarray(txt)]


# --------------------------------------------------------------------------------------
# "separate" quantizer                                            (src/embed.py:187-205)
# --------------------------------------------------------------------------------------
def separate_forward(x, table, asr_w, asr_b, stop_grad=True, phn_attr=None, proj_w=None,
                     proj_b=None, emb_weight=None, dtype=np.float64):
    """p_code = softmax(x @ asr_w.T + asr_b) (:190); idx = argmax (:193);
    stop_grad: new_latent = table[idx] (:194-197);
    else: cat(p_hard @ emb, proj(p_hard @ phn_attr)) (:199-203)."""
    x = np.asarray(x, dtype=dtype)
    B, S, D = x.shape
    N = B * S
    W = np.asarray(asr_w, dtype=dtype)
    K = W.shape[0]
    logits = x.reshape(N, D) @ W.T + np.asarray(asr_b, dtype=dtype)
    p = softmax_rows(logits)
    idx = first_argmax(p)
    e = np.asarray(table, dtype=dtype)
    if stop_grad:
        new_latent = e[idx]
    else:
        onehot = np.zeros_like(p)
        onehot[np.arange(N), idx] = 1
        p_hard = p + (onehot - p)
        emb = np.asarray(emb_weight, dtype=dtype)
        new_latent = p_hard @ emb
        if phn_attr is not None:
            mix = p_hard @ np.asarray(phn_attr, dtype=dtype)
            new_latent = np.concatenate(
                [new_latent, mix @ np.asarray(proj_w, dtype=dtype).T + np.asarray(proj_b, dtype=dtype)], -1)
    return dict(p_code=p.reshape(B, S, K), idx=idx.reshape(B, S),
                new_latent=new_latent.reshape(B, S, -1), logits=logits.reshape(B, S, K))


def separate_backward(x, table, asr_w, p_code, idx, g_p=None, g_q=None, stop_grad=True,
                      dtype=np.float64):
    """Gradients of sum(p_code*g_p)+sum(new_latent*g_q) for the separate quantizer.

    Returns dict(dx, d_asr_w[K,D], d_asr_b[K], dtable[K,D]).  With stop_grad the gather
    route reaches only the table (no straight-through to x, :194-197); without it
    new_latent = p_hard @ table so g_q @ table.T joins the softmax route (:199-203; for the
    attribute columns proj(p_hard @ A) == p_hard @ proj(A) - (sum p_hard - 1) b, and
    sum p_hard == 1, so the full-table form is exact in value and gradient w.r.t. p)."""
    x = np.asarray(x, dtype=dtype)
    B, S, D = x.shape
    N = B * S
    W = np.asarray(asr_w, dtype=dtype)
    K = W.shape[0]
    xf = x.reshape(N, D)
    P = np.asarray(p_code, dtype=dtype).reshape(N, K)
    ix = np.asarray(idx).reshape(N)
    e = np.asarray(table, dtype=dtype)
    Dq = e.shape[1]
    gp = np.zeros((N, K), dtype) if g_p is None else np.asarray(g_p, dtype=dtype).reshape(N, K)
    gq = np.zeros((N, Dq), dtype) if g_q is None else np.asarray(g_q, dtype=dtype).reshape(N, Dq)
    dtable = np.zeros((K, Dq), dtype)
    G = gp.copy()
    if stop_grad:
        np.add.at(dtable, ix, gq)
    else:
        G = G + gq @ e.T
        onehot = np.zeros((N, K), dtype)
        onehot[np.arange(N), ix] = 1
        dtable += onehot.T @ gq
    s = np.sum(G * P, axis=-1, keepdims=True)
    Gs = P * (G - s)
    return dict(dx=(Gs @ W).reshape(B, S, D), d_asr_w=Gs.T @ xf, d_asr_b=np.sum(Gs, axis=0),
                dtable=dtable)


# --------------------------------------------------------------------------------------
# loss extensions (NO reference arithmetic -- parity UNPINNED; van den Oord et al. 2017)
# --------------------------------------------------------------------------------------
def vq_losses(x, code, dtype=np.float64):
    """vq_loss = mean((sg(x) - c)^2), commit_loss = mean((x - sg(c))^2); identical values."""
    x = np.asarray(x, dtype=dtype)
    c = np.asarray(code, dtype=dtype)
    m = np.mean((x - c) ** 2)
    return dict(vq_loss=m, commit_loss=m)


def vq_losses_backward(x, code, idx, K, g_vq=0.0, g_commit=0.0, dtype=np.float64):
    """d vq_loss / d table = scatter_add(idx, 2 (c - x) / (N D)); d commit / d x = 2 (x - c) / (N D)."""
    x = np.asarray(x, dtype=dtype)
    c = np.asarray(code, dtype=dtype)
    D = x.shape[-1]
    xf, cf = x.reshape(-1, D), c.reshape(-1, D)
    n = xf.size
    dx = dtype(g_commit) * 2 * (xf - cf) / n
    dtable = np.zeros((K, D), dtype)
    np.add.at(dtable, np.asarray(idx).reshape(-1), dtype(g_vq) * 2 * (cf - xf) / n)
    return dict(dx=dx.reshape(x.shape), dtable=dtable)


# --------------------------------------------------------------------------------------
# code-usage histogram            (bin/train_vqvae.py:256-261,305,310; src/util.py:135-145)
# --------------------------------------------------------------------------------------
def usage_counts(idx, K):
    """Raw per-code counts of one step's picked indices (what `tok_usage += ...tolist()` accumulates)."""
    return np.bincount(np.asarray(idx).reshape(-1), minlength=K).astype(np.int64)


def usage_bar(counts, zero_pad_tok=True):
    """data_to_bar's `cnts`: counts / total with cnts[0] forced to 0 (src/util.py:139-143)."""
    c = np.asarray(counts, dtype=np.float64)
    tot = c.sum()
    out = c / tot if tot > 0 else c
    if zero_pad_tok:
        out = out.copy()
        out[0] = 0
    return out


# --------------------------------------------------------------------------------------
# run-length collapse after the quantizer                         (src/vqvae.py:218-257)
# --------------------------------------------------------------------------------------
def mean_forward(idx, latent, max_frames_per_phn):
    """idx[B,T] int, latent[B,T,D] -> (padded[B,Lmax,D], lens[B]) or None if a sample is all blank.

    A new segment starts at t when idx changes or the current run is longer than
    max_frames_per_phn (:231); blank (0) segments are dropped (:233); the last segment is the
    mean of latent[last_pos:] unless it starts at T-1, in which case it is latent[T-1] (:239-245).
    """
    idx = np.asarray(idx)
    latent = np.asarray(latent)
    B, T, D = latent.shape
    outs, lens = [], []
    for b in range(B):
        seq = idx[b].tolist()
        last_idx, last_pos, cur = seq[0], 0, []
        for t, k in enumerate(seq):
            if last_idx != k or (t - last_pos) > max_frames_per_phn:
                if last_idx != 0:
                    cur.append(latent[b, last_pos:t].mean(axis=0))
                last_idx, last_pos = k, t
        if last_idx != 0:
            if last_pos != T - 1:
                cur.append(latent[b, last_pos:].mean(axis=0))
            else:
                cur.append(latent[b, T - 1])
        if not cur:
            return None
        lens.append(len(cur))
        outs.append(np.stack(cur, 0))
    L = max(lens)
    padded = np.zeros((B, L, D), latent.dtype)
    for b, o in enumerate(outs):
        padded[b, :len(o)] = o
    return padded, np.asarray(lens, dtype=np.int64)


def mean_forward_segments(idx_row, max_frames_per_phn):
    """Kept (non-blank) segments [start, end) of one utterance, by the same scan as mean_forward (:228-245)."""
    seq = list(np.asarray(idx_row).tolist())
    T = len(seq)
    last_idx, last_pos, segs = seq[0], 0, []
    for t, k in enumerate(seq):
        if last_idx != k or (t - last_pos) > max_frames_per_phn:
            if last_idx != 0:
                segs.append((last_pos, t))
            last_idx, last_pos = k, t
    if last_idx != 0:
        segs.append((last_pos, T))
    return segs


def mean_forward_backward(idx, g_out, max_frames_per_phn):
    """Autograd of mean_forward w.r.t. latent: every frame of a kept segment receives g_out[b, j] / len(segment);
    blank frames receive 0 (mean(dim=0) backward of :234/:242; the single-frame case :245 is the same formula)."""
    idx = np.asarray(idx)
    g_out = np.asarray(g_out)
    B, T = idx.shape
    d = np.zeros((B, T, g_out.shape[-1]), g_out.dtype)
    for b in range(B):
        for j, (s, e) in enumerate(mean_forward_segments(idx[b], max_frames_per_phn)):
            d[b, s:e] = g_out[b, j] / (e - s)
    return d


# --------------------------------------------------------------------------------------
# CTC input preparation                              (bin/train_vqvae.py:430-432, :236)
# --------------------------------------------------------------------------------------
CTC_EPS = 1e-10          # EPS of bin/train_vqvae.py:19


def ctc_input(p_code, eps=CTC_EPS, dtype=np.float64):
    """ctc_input = (model_output + EPS).transpose(0, 1).log(): [B,S,K] -> [S,B,K]."""
    p = np.asarray(p_code, dtype=dtype)
    return np.log(p + dtype(eps)).transpose(1, 0, 2)


def ctc_input_backward(p_code, g_out, eps=CTC_EPS, dtype=np.float64):
    """d/dp of the above: g_p[b,s,k] = g_out[s,b,k] / (p[b,s,k] + eps)."""
    p = np.asarray(p_code, dtype=dtype)
    return np.asarray(g_out, dtype=dtype).transpose(1, 0, 2) / (p + dtype(eps))


# --------------------------------------------------------------------------------------
# helpers for parity reports
# --------------------------------------------------------------------------------------
def top2_rel_gap(dist):
    """Relative gap between the two smallest distances of each row: (d2 - d1) / max(|d1|, tiny)."""
    d = np.asarray(dist, dtype=np.float64)
    d = d.reshape(-1, d.shape[-1])
    if d.shape[1] < 2:
        return np.full(d.shape[0], np.inf)
    part = np.partition(d, 1, axis=-1)[:, :2]
    lo, hi = part.min(-1), part.max(-1)
    return (hi - lo) / np.maximum(np.abs(lo), 1e-30)


def index_mismatch_report(idx_a, idx_b, dist64, rel_gap=1e-6):
    """Counts rows where two index sets differ, split into near-ties (exempt) and hard mismatches."""
    a = np.asarray(idx_a).reshape(-1)
    b = np.asarray(idx_b).reshape(-1)
    gap = top2_rel_gap(dist64)
    diff = a != b
    near = gap < rel_gap
    return dict(rows=int(a.size), mismatched=int(diff.sum()), near_tie_rows=int(near.sum()),
                exempt_mismatches=int((diff & near).sum()), hard_mismatches=int((diff & ~near).sum()))
