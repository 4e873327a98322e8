"""Load tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified reference)."""
import os

import numpy as np

from tencent_recommendation_2025_b200.layout import FeatureLayout, DEFAULT_FEAT_TYPES
from tencent_recommendation_2025_b200.synth import PackedCall

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAIN_CASES = ["baseline_h32", "o1_h64", "o1_h64_mm2", "baseline_l102_nomm"]


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        z = self.z
        self.B, self.L, self.H = int(z["B"]), int(z["L"]), int(z["H"])
        self.mm_ids = [str(x) for x in z["mm_ids"]]
        stats = {str(k): int(v) for k, v in zip(z["stat_keys"], z["stat_vals"])}
        ft = {k: list(v) for k, v in DEFAULT_FEAT_TYPES.items()}
        ft["item_emb"] = self.mm_ids
        self.feat_types, self.feat_statistics = ft, stats
        self.item_num, self.user_num = int(z["item_num"]), int(z["user_num"])
        self.layout = FeatureLayout(self.user_num, self.item_num, stats, ft, self.H)
        self.n_steps = int(z["n_steps"]) if "n_steps" in z else 0
        self.lr = float(z["lr"]) if "lr" in z else None
        self.wd = float(z["wd"]) if "wd" in z else None

    def params0(self):
        return {k[len("param0/"):]: self.z[k] for k in self.z.files if k.startswith("param0/")}

    def group(self, prefix):
        return {k[len(prefix):]: self.z[k] for k in self.z.files if k.startswith(prefix)}

    def call(self, step, c) -> PackedCall:
        pre = f"s{step}/c{c}/"
        z = self.z
        mm = [z[pre + f"mm_x{j}"] for j in range(len(self.mm_ids))]
        inc = (pre + "mask") in z.files
        return PackedCall(self.B, self.L, inc, z[pre + "ids"], z[pre + "arr_off"], z[pre + "arr_val"], mm,
                          seq=z[pre + "seq"], mask=z[pre + "mask"] if inc else None)

    def calls(self, step):
        return [self.call(step, c) for c in range(3)]

    def outs(self, step):
        return [self.z[f"s{step}/c{c}/out"] for c in range(3)]

    def upstream(self, step):
        return [self.z[f"s{step}/c{c}/upstream"] for c in range(3)]

    def sweep_call(self) -> PackedCall:
        z = self.z
        mm = [z[f"mm_x{j}"] for j in range(len(self.mm_ids))]
        return PackedCall(1, self.L, False, z["ids"], z["arr_off"], z["arr_val"], mm, seq=z["seq"], mask=None)


def assert_rows_updated(got, want, grad, lr, rtol=1e-5, what=""):
    """Updated table rows at the north_star bar (1e-5 of tensor scale) with an explicit |g| guard. AdamW's first step moves
    an element by lr * g / (|g| + eps): where |g| is below the gradient's own resolution (rtol of the gradient scale) ANY
    fp32 summation order — the reference's included — lands anywhere within the 2 * lr a sign flip costs, so those
    elements are held to 2 * lr and everything else to rtol. The guard may excuse at most 1 % of the elements."""
    import numpy as np
    got, want, grad = np.asarray(got, np.float64), np.asarray(want, np.float64), np.asarray(grad, np.float64)
    assert got.shape == want.shape == grad.shape, (what, got.shape, want.shape, grad.shape)
    if got.size == 0:
        return
    d = np.abs(got - want)
    wscale, gscale = max(np.abs(want).max(), 1e-30), max(np.abs(grad).max(), 1e-30)
    firm = np.abs(grad) >= rtol * gscale
    assert d[firm].max(initial=0.0) <= rtol * wscale, f"{what}: updated rows off by {d[firm].max():.3e} (scale {wscale:.3e})"
    assert d[~firm].max(initial=0.0) <= 2.0 * lr * 1.001 + rtol * wscale, f"{what}: sub-resolution gradients moved {d[~firm].max():.3e}"
    assert firm.mean() >= 0.99, f"{what}: guard excuses {1 - firm.mean():.4f} of the elements"
