"""TEST INFRASTRUCTURE: a numpy implementation of the ShardOps interface (sharded.CudaShardOps) built on the
CPU oracle, so the exchange orchestration (generators, split sizes, reverse routing, owner-side merge) can run
under a real gloo process group without a GPU. Never imported by the product."""
import numpy as np
import torch

from oracle import feat2emb_numpy as onp
from tencent_recommendation_2025_b200.synth import PackedCall


def _pc(pb) -> PackedCall:
    return PackedCall(pb.B, pb.L, pb.include_user, pb.ids.numpy(), pb.arr_off.numpy(), pb.arr_val.numpy(),
                      [x.numpy() for x in pb.mm_x])


class NumpyShardOps:
    def __init__(self, layout, local_table: torch.Tensor, mm_params, W):
        self.layout, self.W = layout, W
        self.local = local_table
        self.m = np.zeros_like(local_table.numpy())
        self.v = np.zeros_like(local_table.numpy())
        self.mm_params = mm_params          # {fid: (W [H, d], b [H])} numpy
        self.step = 0

    def unique_keys(self, pbs):
        keys, _ = onp.build_keys(self.layout, [_pc(pb) for pb in pbs])
        uniq = np.unique(keys).astype(np.uint32)
        cap = max(keys.size, 1)
        out = np.zeros(cap, np.int32)
        out[:uniq.size] = uniq.view(np.int32)
        return torch.from_numpy(out), torch.tensor([uniq.size], dtype=torch.int32), cap

    def route(self, uniq, n_unique, cap):
        n = int(n_unique[0])
        u = uniq.numpy()[:n].view(np.uint32)
        owner, local, counts, order = onp.route(u, self.W)
        rows = np.zeros(cap, np.int32)
        rows[:n] = local[order].astype(np.int32)
        perm = np.zeros(cap, np.int32)
        perm[order] = np.arange(n, dtype=np.int32)
        return torch.from_numpy(rows), torch.from_numpy(perm), torch.from_numpy(counts.astype(np.int32))

    def gather(self, rows, n):
        return self.local[rows[:n].long()].clone()

    def forward_from_rows(self, pb, uniq, n_unique, perm, rows_buf, out_dtype):
        lay = self.layout
        n = int(n_unique[0])
        u = uniq.numpy()[:n].view(np.uint32).astype(np.int64)
        pm = perm.numpy()
        buf = rows_buf.numpy()
        pc = _pc(pb)
        cl = lay.calls[pb.include_user]
        T, H = pb.T, lay.H
        item = np.zeros((T, cl.item_dim), np.float32)
        user = np.zeros((T, cl.user_dim), np.float32) if pb.include_user else None

        def rows_of(ids, table):
            keys = lay.tables[table].key_base + ids.astype(np.int64)
            idx = np.searchsorted(u, keys)
            idx = np.minimum(idx, max(n - 1, 0))
            r = np.where(ids != 0, 1 + pm[idx], 0)
            return buf[r]

        for s in cl.slots:
            dst = item if s.side == 0 else user
            if s.kind == 0:
                dst[:, s.col:s.col + H] = rows_of(pc.ids[:, s.src], s.table)
            elif s.kind == 1:
                off = pc.arr_off[s.src].astype(np.int64)
                for t in np.nonzero(np.diff(off))[0]:
                    acc = np.zeros(H, np.float32)
                    for r in rows_of(pc.arr_val[off[t]:off[t + 1]], s.table):
                        acc = (acc + r).astype(np.float32)
                    dst[t, s.col:s.col + H] = acc
            else:
                w, b = self.mm_params[s.name]
                dst[:, s.col:s.col + H] = (pc.mm_x[s.src] @ w.T + b).astype(np.float32)
        return torch.from_numpy(item), None if user is None else torch.from_numpy(user)

    def reduce(self, calls):
        pcs = [_pc(pb) for pb, _, _ in calls]
        d = [(di.numpy(), None if du is None else du.numpy()) for _, di, du in calls]
        uniq, rows = onp.segment_reduce_fp64(self.layout, pcs, d)
        cap = max(sum(int(np.count_nonzero(p.ids)) + p.arr_val.size for p in pcs), 1)
        u = np.zeros(cap, np.int32)
        u[:uniq.size] = uniq.astype(np.uint32).view(np.int32)
        g = np.zeros((cap, self.layout.H), np.float32)
        g[:uniq.size] = rows.astype(np.float32)
        return torch.from_numpy(u), torch.tensor([uniq.size], dtype=torch.int32), torch.from_numpy(g), cap

    def permute(self, grads, perm, n_unique, cap):
        n = int(n_unique[0])
        out = torch.zeros_like(grads)
        out[perm[:n].long()] = grads[:n]
        return out

    def apply(self, recv_rows, recv_grads, R, hyper):
        self.step += 1
        if R == 0:
            return
        rows = recv_rows.numpy()[:R].astype(np.int64)
        g = recv_grads.numpy()[:R]
        order = np.argsort(rows, kind="stable")
        rs, gs = rows[order], g[order]
        uniq, start = np.unique(rs, return_index=True)
        seg = np.append(start, R)
        red = np.zeros((uniq.size, g.shape[1]), np.float32)
        for i in range(uniq.size):
            acc = np.zeros(g.shape[1], np.float32)
            for j in range(seg[i], seg[i + 1]):
                acc = (acc + gs[j]).astype(np.float32)      # source-rank order inside a row
            red[i] = acc
        w = self.local.numpy()
        onp.adamw_rows(w, self.m, self.v, uniq, red, self.step, lr=hyper["lr"], beta1=hyper["betas"][0],
                       beta2=hyper["betas"][1], eps=hyper["eps"], wd=hyper["weight_decay"])

    # ---- step-level prefetch protocol (mirrors CudaShardOps.prepare / reduce_cached / prepare_owner / apply_cached)
    def prepare(self, pbs):
        pcs = [_pc(pb) for pb in pbs]
        keys, srcs = onp.build_keys(self.layout, pcs)
        uniq = np.unique(keys).astype(np.uint32)
        cap = max(keys.size, 1)
        out = np.zeros(cap, np.int32)
        out[:uniq.size] = uniq.view(np.int32)
        return {"pcs": pcs, "n": keys.size, "uniq": torch.from_numpy(out),
                "n_unique": torch.tensor([uniq.size], dtype=torch.int32), "cap": cap}

    def reduce_cached(self, pf, calls):
        d = [(di.numpy(), None if du is None else du.numpy()) for _, di, du in calls]
        uniq, rows = onp.segment_reduce_fp64(self.layout, pf["pcs"], d)
        g = np.zeros((pf["cap"], self.layout.H), np.float32)
        g[:uniq.size] = rows.astype(np.float32)
        return torch.from_numpy(g)

    def prepare_owner(self, recv_rows, R, recv_counts=None):
        return recv_rows.clone()

    def apply_cached(self, owner_state, recv_grads, R, hyper):
        self.apply(owner_state, recv_grads, R, hyper)


class NumpyPeerShardOps(NumpyShardOps):
    """The peer-memory flavour of the protocol (sharded.FactShardOps) with numpy arrays standing in for NVLink peer
    memory: ranks emulated in ONE process see each other's shard / gradient window objects directly. Rows are read in
    place from their owners (no row all-to-all), the owner pulls gradient rows out of every source's window."""

    def __init__(self, layout, local_table, mm_params, W, window_rows):
        super().__init__(layout, local_table, mm_params, W)
        self.peers = None
        self.grad_peers = None
        self.grad_win = torch.zeros((window_rows, layout.H))
        self._bar = torch.zeros(1)
        self.pulled = 0          # steps whose gradients were pulled through the window

    def link(self, all_ops):
        self.peers = list(all_ops)
        self.grad_peers = list(all_ops)

    def fetch_rows_async(self, pf):
        n = int(pf["n_unique"][0])
        u = pf["uniq"].numpy()[:n].view(np.uint32).astype(np.int64)
        rows = np.zeros((n + 1, self.layout.H), np.float32)          # row 0 = the zero row the concat builder expects
        for i, k in enumerate(u):
            rows[1 + i] = self.peers[int(k % self.W)].local.numpy()[int(k // self.W)]
        pf["rows_sorted"] = torch.from_numpy(rows)

    def forward_prefetched(self, pb, st, out_dtype=None):
        pf = st["pf"]
        ident = torch.arange(pf["cap"], dtype=torch.int32)
        return self.forward_from_rows(pb, pf["uniq"], pf["n_unique"], ident, pf["rows_sorted"], out_dtype)

    def prepare_owner(self, recv_rows, R, counts_dev=None, recv_counts=None):
        if counts_dev is None:
            return super().prepare_owner(recv_rows, R)
        cnt = counts_dev.numpy().astype(np.int64)
        src = np.repeat(np.arange(self.W), cnt)
        idx = np.arange(R) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        return recv_rows.clone(), src, idx

    def permute_to_window(self, grads, perm, n_unique, cap):
        n = int(n_unique[0])
        self.grad_win[perm[:n].long()] = grads[:n]

    def apply_from_peers(self, owner_state, starts, R, hyper):
        rows, src, idx = owner_state
        g = np.zeros((max(R, 1), self.layout.H), np.float32)
        for j in range(R):
            g[j] = self.grad_peers[int(src[j])].grad_win.numpy()[starts[int(src[j])] + int(idx[j])]
        self.pulled += 1
        self.apply(rows, torch.from_numpy(g), R, hyper)
