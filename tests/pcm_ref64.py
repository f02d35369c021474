"""fp64 companions of the PCM parity tests (test infrastructure, runs on the GPU in float64).

The PCM gradient contains [S_ab > 0] (the derivative of the relu in network/resnet38_contrast.py:71), a step function of
a quantity that passes through zero: where |S_ab| is at the level of fp32 rounding (~1e-7) two correct fp32 evaluations —
the oracle's cuBLAS product and ours — may decide the sign differently, and each decision adds or drops a whole term of
the gradient.  The helpers here (a) evaluate S in fp64 and list those AMBIGUOUS entries, (b) read the sign decisions each
side actually took, (c) evaluate the gradient terms of exactly those entries, so a test can take them out of BOTH
gradients and hold everything else to the max-abs bar.
"""
import torch


def decode_mask(ws, lay, n, hw):
    """The sign mask a PCM forward left in its workspace -> bool [n, HW(j), HW(i)] (True = S_ji counted as positive).
    Layout: mask[n][stage][row j][4 words]; in a word, bit m = column 2m of the 32-column chunk, bit 16+m = column 2m+1."""
    HWpj, HWpi, NS, wb = lay["HWpj"], lay["HWpi"], lay["NS"], lay["wb"]
    raw = ws[lay["mask"]:lay["mask"] + n * NS * HWpj * 16].view(torch.int32).reshape(n, NS, HWpj, 4)
    order = torch.stack([torch.arange(16), torch.arange(16) + 16], 1).reshape(-1).to(ws.device).int()
    out = torch.empty(n, hw, HWpi, dtype=torch.bool, device=ws.device)
    for s in range(NS):                                  # stage by stage: the unpacked bits of all stages at once are large
        bits = ((raw[:, s, :hw].unsqueeze(-1) >> order) & 1).reshape(n, hw, 128)[..., :2 * wb]
        out[:, :, s * 2 * wb:(s + 1) * 2 * wb] = bits.bool()
    return out[:, :, :hw]


class PcmRef64:
    """Everything the backward formula needs, in fp64:
    dL/dA_ab = cam_a . g_b - d_b with g = G / D, d_b = sum_c out[c,b] G[c,b] / D_b;  dS = dA * [S > 0];
    dfh_a = sum_b (dS_ab + dS_ba) fh_b;  df = dfh / (|f| + eps) - f (dfh . f) / (|f| (|f| + eps)^2)."""

    def __init__(self, camn, f, gout):
        n, k, h, w = f.shape
        hw = h * w
        self.shape = (n, k, h, w)
        self.f = f.reshape(n, k, hw).double()
        self.nrm = self.f.norm(dim=1, keepdim=True)
        self.fh = (self.f / (self.nrm + 1e-5)).transpose(1, 2).contiguous()          # [n, HW, K]
        cam = camn.reshape(n, -1, hw).double()
        G = gout.reshape(n, -1, hw).double()
        self.S = torch.matmul(self.fh, self.fh.transpose(1, 2))                      # [n, HW, HW]
        A = self.S.clamp_min(0)
        D = A.sum(dim=1) + 1e-5                                                      # [n, HW] column sums
        out = torch.matmul(cam, A) / D.unsqueeze(1)
        self.cam = cam.transpose(1, 2).contiguous()                                  # [n, HW, C]
        self.g = (G / D.unsqueeze(1)).transpose(1, 2).contiguous()                   # [n, HW, C]
        self.d = (out * G).sum(dim=1) / D                                            # [n, HW]

    def ambiguous(self, thresh=1e-6):
        """Index triple (image, a, b) of the entries with |S_ab| < thresh, and their share of all entries."""
        amb = self.S.abs() < thresh
        idx = amb.nonzero(as_tuple=True)
        return idx, idx[0].numel() / amb.numel()

    def terms(self, idx, m_ab, m_ba):
        """Gradient w.r.t. f of ONLY the listed entries: for the row a of each (image, a, b), the term
        (m_ab * dA_ab + m_ba * dA_ba) fh_b — m_ab / m_ba are the sign decisions taken for S_ab / S_ba."""
        n, k, h, w = self.shape
        ni, a, b = idx
        dA_ab = (self.cam[ni, a] * self.g[ni, b]).sum(-1) - self.d[ni, b]
        dA_ba = (self.cam[ni, b] * self.g[ni, a]).sum(-1) - self.d[ni, a]
        coef = m_ab.double() * dA_ab + m_ba.double() * dA_ba
        dfh = torch.zeros_like(self.fh)
        dfh.index_put_((ni, a), coef.unsqueeze(-1) * self.fh[ni, b], accumulate=True)
        dfh = dfh.transpose(1, 2)                                                    # [n, K, HW]
        dot = (dfh * self.f).sum(dim=1, keepdim=True)
        nrm = self.nrm
        df = dfh / (nrm + 1e-5) - self.f * dot / (nrm.clamp_min(1e-30) * (nrm + 1e-5) ** 2)
        return df.reshape(n, k, h, w)
