"""CPU emulation of the clipk kernel entry points, for tests of the HOST logic only (collectives, mode coefficients,
autograd plumbing) under gloo.  Same method signatures and return conventions as clipk.ops.CudaBackend; plain torch
float64 arithmetic.  Never used by the product: clipk.ops._backend() only returns it after a test called
set_backend_for_testing()."""
import torch


class EmuOperand:
    def __init__(self, x):
        self.data = x.detach().to(torch.float64)
        self.rows, self.d = x.shape
        self.dtype = -1


class EmuBackend:
    launches = 0

    def prepare(self, x):
        return EmuOperand(x)

    def prepare_grad(self, op):
        return op

    def fwd_stats(self, X, Y, scale, diag_offset, want_pos, out=None):
        S = float(scale[0]) * X.data @ Y.data.T
        m = S.max(dim=1).values
        e = torch.exp(S - m[:, None])
        stats = torch.stack((m, e.sum(1), (e * S).sum(1))).to(torch.float32)
        if out is not None:
            out.copy_(stats)
            stats = out
        pos = None
        if want_pos:
            idx = torch.arange(X.rows) + diag_offset
            pos = S[torch.arange(X.rows), idx].to(torch.float32)
        return stats, pos

    def fwd_both(self, X, Y, scale, diag_offset, col_out=None, exact=False):
        row_stats, pos = self.fwd_stats(X, Y, scale, diag_offset, True)
        col_stats, _ = self.fwd_stats(Y, X, scale, 0, False, out=col_out)
        return row_stats, pos, col_stats

    def finalize(self, row_stats, pos, col_parts, diag_offset):
        rs = row_stats.double()
        cp = col_parts.double()
        lse_row = rs[0] + rs[1].log()
        M = cp[:, 0].max(dim=0).values
        w = torch.exp(cp[:, 0] - M)
        L = (cp[:, 1] * w).sum(0)
        Tt = (cp[:, 2] * w).sum(0)
        lse_col = M + L.log()
        e_col = Tt / L
        rows = rs.shape[1]
        idx = torch.arange(rows) + diag_offset
        p = pos.double()
        sums = torch.stack(((lse_row - p).sum(), (lse_col[idx] - p).sum(), (rs[2] / rs[1] - p).sum(),
                            (e_col[idx] - p).sum()))
        return lse_row.float(), lse_col.float(), sums.float()

    def bwd(self, X, Y, Xg, Yg, scale, diag_offset, lse_row, lse_col, alpha, beta, gscale, want_dx, want_dy, split=False):
        s = float(scale[0])
        S = s * X.data @ Y.data.T
        Pr = torch.exp(S - lse_row.double()[:, None])
        Pc = torch.exp(S - lse_col.double()[None, :])
        eye = torch.zeros_like(S)
        eye[torch.arange(X.rows), torch.arange(X.rows) + diag_offset] = 1.0
        k = s * float(gscale[0])
        Gx = k * alpha * (Pr - eye) if split else k * (alpha * (Pr - eye) + beta * (Pc - eye))
        Gy = k * beta * (Pc - eye) if split else Gx
        dX = (Gx @ Y.data).float() if want_dx else None
        dY = (Gy.T @ X.data).float() if want_dy else None
        return dX, dY

    # ---- the fused step (clipk_step_forward / clipk_step_backward); the peer-memory collectives become gloo ones
    class _Peer:
        pass

    def peer_context(self, b, d, rank, world, group):
        return self._Peer()

    def step_workspace_bytes(self, rows, cols, d, world):
        return 256

    def step_forward(self, st):
        import torch.distributed as dist
        W, b, N, rank = st.world, st.rows, st.cols, st.rank
        x, y = st.image.double(), st.text.double()
        if st.normalize:
            ix = 1.0 / x.norm(dim=-1).clamp_min(st.eps)
            iy = 1.0 / y.norm(dim=-1).clamp_min(st.eps)
            st.inv_x.copy_(ix.float())
            st.inv_y.copy_(iy.float())
            x, y = x * ix[:, None], y * iy[:, None]
        xb, yb = x.to(torch.bfloat16), y.to(torch.bfloat16)
        if st.x_op is not st.image:
            st.x_op.copy_(xb)
        if W > 1:
            parts = [torch.empty(b, st.d, dtype=torch.float32) for _ in range(W)]
            dist.all_gather(parts, yb.float(), group=st.group)
            st.y_all.copy_(torch.cat(parts).to(torch.bfloat16))
        elif st.y_all is not st.text:
            st.y_all.copy_(yb)
        X, Y = st.x_op.double(), st.y_all.double()
        s, off = float(st.scale[0]), rank * b
        S = s * X @ Y.T
        idx = torch.arange(b)
        pos = S[idx, idx + off]
        lse_row = torch.logsumexp(S, dim=1)
        e_row = (torch.softmax(S, dim=1) * S).sum(1)
        cm = S.max(dim=0).values
        ce = torch.exp(S - cm[None, :])
        col = torch.stack((cm, ce.sum(0), (ce * S).sum(0)))
        if W > 1:
            cols = [torch.empty_like(col) for _ in range(W)]
            dist.all_gather(cols, col, group=st.group)
            cp = torch.stack(cols)
        else:
            cp = col[None]
        M = cp[:, 0].max(dim=0).values
        w = torch.exp(cp[:, 0] - M)
        L, Tt = (cp[:, 1] * w).sum(0), (cp[:, 2] * w).sum(0)
        lse_col, e_col = M + L.log(), Tt / L
        st.lse_row.copy_(lse_row.float())
        st.lse_col.copy_(lse_col.float())
        sums = ((lse_row - pos).sum(), (lse_col[idx + off] - pos).sum(), (e_row - pos).sum(), (e_col[idx + off] - pos).sum())
        st.scal[4] = float(sums[0] + sums[1]) / st.loss_div
        st.scal[5] = float(sums[2] + sums[3]) / st.loss_div

    def step_backward(self, st):
        import torch.distributed as dist
        W, b, rank = st.world, st.rows, st.rank
        s, go, off = float(st.scale[0]), float(st.grad_out[0]), rank * b
        if st.d_scale is not None:
            st.d_scale[0] = float(st.scal[5]) * go / s
        if st.d_image is None:
            return
        X, Y = st.x_op.double(), st.y_all.double()
        S = s * X @ Y.T
        Pr, Pc = torch.exp(S - st.lse_row.double()[:, None]), torch.exp(S - st.lse_col.double()[None, :])
        Pr[torch.arange(b), torch.arange(b) + off] -= 1.0
        Pc[torch.arange(b), torch.arange(b) + off] -= 1.0
        k = s * go * st.grad_coef
        Gx, Gy = (k * Pr, k * Pc) if st.grad_split else (k * (Pr + Pc),) * 2
        dX, dY = Gx @ Y, Gy.T @ X
        if W > 1:
            dist.all_reduce(dY, group=st.group)
        dT = dY[off:off + b]
        if st.normalize:
            for g, yn, inv in ((dX, X, st.inv_x.double()), (dT, Y[off:off + b], st.inv_y.double())):
                g.copy_((g - yn * (g * yn).sum(-1, keepdim=True)) * inv[:, None])
        st.d_image.copy_(dX.to(st.d_image.dtype))
        st.d_text.copy_(dT.to(st.d_text.dtype))

    def normalize_fwd(self, x, eps):
        n = x.double().norm(dim=-1, keepdim=True).clamp_min(eps)
        return (x.double() / n).to(x.dtype), (1.0 / n.squeeze(-1)).float()

    def normalize_bwd(self, g, y, inv, eps):
        g64, y64 = g.double(), y.double()
        dot = (g64 * y64).sum(-1, keepdim=True)
        return ((g64 - y64 * dot) * inv.double()[:, None]).to(y.dtype)

    def cast(self, src, dtype):
        return src.to(dtype)

    # evaluation side
    def dense_operand(self, x):
        op = EmuOperand(x)
        op.f16, op.k, op.planes = 0, x.shape[1], [x]
        return op

    # distillation term
    def gemm(self, A, B, out, a_mn, b_mn, f16, accumulate):
        a = A.double().T if a_mn else A.double()
        b = B.double() if b_mn else B.double().T
        prod = (a @ b).float()
        if accumulate:
            out += prod
        else:
            out.copy_(prod)

    def grad_operand(self, x):
        return x.double(), torch.ones(1)

    def distill_cross(self, S, T, nrows, cols, s_mul, t_mul, t_lse_row, t_lse_col, row0, row_cross, col_part):
        s = S[:nrows, :cols].double() * float(s_mul[0])
        t = T[:nrows, :cols].double() * float(t_mul[0])
        row_cross[row0:row0 + nrows] = (torch.exp(t - t_lse_row[row0:row0 + nrows].double()[:, None]) * s).sum(1).float()
        col_part.copy_((torch.exp(t - t_lse_col.double()[None, :cols]) * s).sum(0).float())

    def distill_grad(self, S, T, nrows, cols, s_mul, t_mul, s_lse_row, t_lse_row, s_lse_col, t_lse_col, row0, G):
        s = S[:nrows, :cols].double() * float(s_mul[0])
        t = T[:nrows, :cols].double() * float(t_mul[0])
        rows = slice(row0, row0 + nrows)
        v = (torch.exp(s - s_lse_row[rows].double()[:, None]) - torch.exp(t - t_lse_row[rows].double()[:, None]) +
             torch.exp(s - s_lse_col.double()[None, :cols]) - torch.exp(t - t_lse_col.double()[None, :cols]))
        G[:nrows, :cols] = (16384.0 * v).to(G.dtype)

    def logits_panel(self, Q, K, r0, nrows, out):
        out.zero_()
        assert K.rows == out.shape[1], "keys must be padded to the panel width"
        out[:nrows] = (Q.data[r0:r0 + nrows] @ K.data.T).float()

    def rank_count(self, S, nrows, cols, target, diag_offset, row0, greater, ties):
        rows = torch.arange(row0, row0 + nrows)
        t = target[rows] if target is not None else rows + diag_offset
        ok = (t >= 0) & (t < cols)
        tc = t.clamp(0, cols - 1)
        logits = S[:nrows, :cols]
        thr = logits[torch.arange(nrows), tc][:, None]
        g = (logits > thr).sum(1)
        before = torch.arange(cols)[None, :] < tc[:, None]
        tb = ((logits == thr) & before).sum(1)
        greater[rows] = torch.where(ok, g, torch.full_like(g, -1)).to(torch.int32)
        ties[rows] = torch.where(ok, tb, torch.zeros_like(tb)).to(torch.int32)
