"""Feature-id / table / concat-column layout of the TencentGR sparse-feature embedding path.

Pure Python, no torch: this is the single source of truth that the CUDA path, the
tensorizer, the oracle tests and the multi-GPU router all read.

Follows the reference's declarations (paths relative to /root/reference):
  * table set and ModuleDict insertion order ......... model/BaseLine/model.py:115-116,158-165
  * feature-id -> type grouping, mm dims .............. model/BaseLine/model.py:169-184
  * concat widths (userdim / itemdim) ................. model/BaseLine/model.py:129-139
  * concat column order (append order in feat2emb) .... model/BaseLine/model.py:244-245,252-263,281-299
  * default feature-id lists .......................... model/BaseLine/dataset.py:191-212
BaseLineO1 is identical on all of the above (model/BaseLineO1/model.py:196-280,327-416).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

# model/BaseLine/model.py:183
EMB_SHAPE_DICT = {"81": 32, "82": 1024, "83": 3584, "84": 4096, "85": 3584, "86": 3584}

# model/BaseLine/dataset.py:191-212 (mm ids come from --mm_emb_id, default ['81'])
DEFAULT_FEAT_TYPES = {
    "user_sparse": ["103", "104", "105", "109"],
    "item_sparse": ["100", "117", "111", "118", "101", "102", "119", "120", "114", "112", "121", "115", "122", "116"],
    "item_array": [],
    "user_array": ["106", "107", "108", "110"],
    "item_emb": ["81"],
    "user_continual": [],
    "item_continual": [],
}

# slot kinds (shared with include/tgr_embed.h)
KIND_SINGLE = 0  # one id per token  -> one table row copied into the concat slot
KIND_ARRAY = 1   # ragged id list    -> rows sum-pooled left to right into the concat slot
KIND_MM = 2      # dense mm vector   -> x @ W^T + b written into the concat slot

SIDE_ITEM = 0
SIDE_USER = 1


@dataclass(frozen=True)
class Table:
    """One ``nn.Embedding(rows, H, padding_idx=0)``; ``rows`` already includes the +1 padding row."""

    name: str          # state_dict prefix: 'item_emb', 'user_emb', 'sparse_emb.<fid>'
    rows: int          # vocab + 1
    key_base: int      # first global key of this table (global key = key_base + id)


@dataclass(frozen=True)
class Slot:
    """One H-wide column block of a concat buffer."""

    name: str          # 'item_id', 'user_id', or the feature id string
    kind: int          # KIND_*
    side: int          # SIDE_ITEM / SIDE_USER
    col: int           # first column (elements) inside that side's concat buffer
    table: int         # index into FeatureLayout.tables (KIND_SINGLE / KIND_ARRAY), else -1
    src: int           # KIND_SINGLE: column of the packed ids matrix; KIND_ARRAY: array index; KIND_MM: mm index
    mm_dim: int = 0    # KIND_MM only


@dataclass
class CallLayout:
    """Slots active in one feat2emb call flavour (include_user True / False)."""

    include_user: bool
    slots: List[Slot]
    n_single: int      # columns of the packed ids matrix
    n_array: int
    n_mm: int
    item_dim: int
    user_dim: int      # 0 when include_user is False


class FeatureLayout:
    """Tables, global keys and concat columns for one (feat_statistics, feat_types, H)."""

    def __init__(self, user_num: int, item_num: int, feat_statistics: Dict[str, int],
                 feat_types: Dict[str, Sequence[str]], hidden_units: int):
        if feat_types.get("user_continual") or feat_types.get("item_continual"):
            # model/BaseLine/model.py:278-279 feeds an int64 [B,L,1] into torch.cat with floats; both
            # shipped variants leave these lists empty (dataset.py:211-212). Width 0 only.
            raise NotImplementedError("continual features are not exercised by the reference configs")
        self.user_num = int(user_num)
        self.item_num = int(item_num)
        self.H = int(hidden_units)
        if self.H % 4 != 0:
            raise ValueError("hidden_units must be a multiple of 4 (128-bit row accesses)")
        self.feat_types = {k: list(v) for k, v in feat_types.items()}
        self.user_sparse = {k: int(feat_statistics[k]) for k in feat_types["user_sparse"]}
        self.item_sparse = {k: int(feat_statistics[k]) for k in feat_types["item_sparse"]}
        self.user_array = {k: int(feat_statistics[k]) for k in feat_types["user_array"]}
        self.item_array = {k: int(feat_statistics[k]) for k in feat_types["item_array"]}
        self.item_emb_feat = {k: EMB_SHAPE_DICT[k] for k in feat_types["item_emb"]}

        # ---- tables, in the reference's ModuleDict insertion order (model.py:158-165)
        names_rows: List[Tuple[str, int]] = [("item_emb", self.item_num + 1), ("user_emb", self.user_num + 1)]
        for group in (self.user_sparse, self.item_sparse, self.item_array, self.user_array):
            for k, vocab in group.items():
                names_rows.append((f"sparse_emb.{k}", vocab + 1))
        self.tables: List[Table] = []
        base = 0
        for name, rows in names_rows:
            self.tables.append(Table(name, rows, base))
            base += rows
        self.total_rows = base
        if self.total_rows >= (1 << 32) - 1:
            raise ValueError("global key space must fit 32 bits")
        self.key_bits = max(1, int(self.total_rows - 1).bit_length())
        self._table_index = {t.name: i for i, t in enumerate(self.tables)}

        # ---- concat widths (model.py:129-139)
        H = self.H
        self.item_dim = H * (len(self.item_sparse) + 1 + len(self.item_array)) + H * len(self.item_emb_feat)
        self.user_dim = H * (len(self.user_sparse) + 1 + len(self.user_array))

        self.calls = {True: self._build_call(True), False: self._build_call(False)}

    # ------------------------------------------------------------------
    def table_index(self, name: str) -> int:
        return self._table_index[name]

    def sparse_table_index(self, fid: str) -> int:
        return self._table_index[f"sparse_emb.{fid}"]

    def _build_call(self, include_user: bool) -> CallLayout:
        H = self.H
        slots: List[Slot] = []
        n_single = n_array = n_mm = 0
        col = 0
        # item side: [item_emb | item_sparse... | item_array... | mm...]  (model.py:244,252-254,281-299)
        slots.append(Slot("item_id", KIND_SINGLE, SIDE_ITEM, col, self.table_index("item_emb"), n_single)); n_single += 1; col += H
        for k in self.item_sparse:
            slots.append(Slot(k, KIND_SINGLE, SIDE_ITEM, col, self.sparse_table_index(k), n_single)); n_single += 1; col += H
        for k in self.item_array:
            slots.append(Slot(k, KIND_ARRAY, SIDE_ITEM, col, self.sparse_table_index(k), n_array)); n_array += 1; col += H
        for k, d in self.item_emb_feat.items():
            slots.append(Slot(k, KIND_MM, SIDE_ITEM, col, -1, n_mm, d)); n_mm += 1; col += H
        assert col == self.item_dim
        ucol = 0
        if include_user:
            # user side: [user_emb | user_sparse... | user_array...]  (model.py:245,260-262)
            slots.append(Slot("user_id", KIND_SINGLE, SIDE_USER, ucol, self.table_index("user_emb"), n_single)); n_single += 1; ucol += H
            for k in self.user_sparse:
                slots.append(Slot(k, KIND_SINGLE, SIDE_USER, ucol, self.sparse_table_index(k), n_single)); n_single += 1; ucol += H
            for k in self.user_array:
                slots.append(Slot(k, KIND_ARRAY, SIDE_USER, ucol, self.sparse_table_index(k), n_array)); n_array += 1; ucol += H
            assert ucol == self.user_dim
        return CallLayout(include_user, slots, n_single, n_array, n_mm, self.item_dim, ucol)

    # ------------------------------------------------------------------
    def single_slot_names(self, include_user: bool) -> List[str]:
        return [s.name for s in self.calls[include_user].slots if s.kind == KIND_SINGLE]

    def array_slot_names(self, include_user: bool) -> List[str]:
        return [s.name for s in self.calls[include_user].slots if s.kind == KIND_ARRAY]

    def mm_slot_names(self) -> List[str]:
        return list(self.item_emb_feat.keys())

    def describe(self) -> str:
        lines = [f"H={self.H} item_dim={self.item_dim} user_dim={self.user_dim} total_rows={self.total_rows} key_bits={self.key_bits}"]
        for i, t in enumerate(self.tables):
            lines.append(f"  table[{i}] {t.name:<18} rows={t.rows:<10} key_base={t.key_base}")
        return "\n".join(lines)


def default_feat_statistics(item_sparse_vocab: Optional[Sequence[int]] = None) -> Dict[str, int]:
    """Synthetic vocabularies of SURVEY.md §8(d) (the real ones are not shipped with the reference)."""
    ft = DEFAULT_FEAT_TYPES
    st: Dict[str, int] = {}
    for i, k in enumerate(ft["item_sparse"]):
        st[k] = int(item_sparse_vocab[i]) if item_sparse_vocab is not None else 10 ** (2 + (i % 5))
    for k, v in zip(ft["user_sparse"], (10, 100, 1000, 10000)):
        st[k] = v
    for k, v in zip(ft["user_array"], (1000, 10000, 100000, 1000)):
        st[k] = v
    return st
