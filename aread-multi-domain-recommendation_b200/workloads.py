"""Synthetic workloads of the shapes BASELINE.json names.  No dataset is read:
vocabulary sizes come from the reference's config / sample files, ids are drawn from seeded
distributions.  Batches are single-domain, like the reference's per-domain loaders
(run.py:326-335, 551-575).
"""
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

# reference config.py:59-64
AMAZON_DOMAIN_SIZE = [69360, 282546, 776105, 3001846, 88496, 449031, 2859592, 1893, 1437340, 16454, 601698, 1802,
                      2416380, 197170, 202176, 6931, 317131, 132650, 602500, 585227, 845268, 1107407, 997451, 623565,
                      44843]
ALICCP_DOMAIN_SIZE = [2695782, 1433175, 925817, 584726, 461755, 358265, 166869, 113621, 78692, 65313, 54483, 45808,
                      40975, 37939, 34079, 31703, 29551, 27084, 25027, 23464, 21764, 19857, 18390, 16712, 15852,
                      14914, 13653, 12265, 11179, 9760]
# field cardinalities of the bundled AliCCP sample, in run.py:57-59 column order
ALICCP_FIELD_DIMS = [211161, 95, 14, 3, 8, 4, 4, 3, 5, 41775, 30, 284915, 81491, 112993, 1929, 118091, 54472, 34677,
                     5821, 106908, 54295, 31716, 4]
# Amazon: itemid, weekday, domain, sales_chart, sales_rank, brand, price (config.py:7 + sample CSV)
AMAZON_FIELD_DIMS = [1368287, 7, 25, 45, 11, 22339, 10]


@dataclass
class Workload:
    name: str
    one_hot_field_dims: Sequence[int]
    domain_idx: int
    itemid_idx: int
    n_domain: int
    domain_size: Sequence[int]
    n_history_fields: int = 0
    seq_maxlen: int = 5
    method: Optional[str] = "mean"
    pad_id: Optional[int] = None
    label_rate: float = 0.5
    embed_dim: int = 32
    id_dist: str = "zipf"

    @property
    def multi_hot_dict(self):
        n_oh = len(self.one_hot_field_dims)
        return {"multi_hot_flag": [False] * n_oh + [True] * (self.n_history_fields * self.seq_maxlen),
                "itemid_idx": self.itemid_idx, "seq_maxlen": self.seq_maxlen, "method": self.method}

    @property
    def n_cols(self):
        return len(self.one_hot_field_dims) + self.n_history_fields * self.seq_maxlen

    @property
    def n_fields(self):
        return len(self.one_hot_field_dims) + self.n_history_fields

    @property
    def n_rows(self):
        return int(np.sum(self.one_hot_field_dims))

    def gather_bytes_per_sample(self):
        """Algorithmic bytes of the lookup: 4 B id + D*4 B row read per looked-up row,
        D*4 B written per output field."""
        row = self.embed_dim * 4
        one_hot = len(self.one_hot_field_dims) * (4 + row + row)
        pooled = self.n_history_fields * (self.seq_maxlen * (4 + row) + row)
        return one_hot + pooled

    def scatter_bytes(self, n_lookups, n_unique):
        return n_lookups * (4 + self.embed_dim * 4) + n_unique * self.embed_dim * 4

    def _ids(self, rng, V, n):
        if self.id_dist == "uniform" or V < 8:
            return rng.randint(0, V, size=n)
        # log-uniform ranks (Zipf exponent ~1), scattered over the field by a multiplicative hash
        rank = np.minimum((V ** rng.random_sample(n)).astype(np.int64) - 1, V - 1)
        return (rank * 2654435761) % V

    def sample_domain(self, rng):
        p = np.asarray(self.domain_size, dtype=np.float64)
        return int(rng.choice(self.n_domain, p=p / p.sum()))

    def batch(self, B, seed, domain=None):
        """(x int32 [B, n_cols], y int16 [B, 1], domain)."""
        rng = np.random.RandomState(seed)
        if domain is None:
            domain = self.sample_domain(rng)
        cols = [self._ids(rng, int(V), B) for V in self.one_hot_field_dims]
        cols[self.domain_idx] = np.full(B, domain, dtype=np.int64)
        V_item = int(self.one_hot_field_dims[self.itemid_idx])
        for _ in range(self.n_history_fields):
            seq = self._ids(rng, V_item, B * self.seq_maxlen).reshape(B, self.seq_maxlen)
            n_real = np.minimum(rng.geometric(1 / 2.7, size=B), self.seq_maxlen)       # mean history ~2.7
            pad = np.arange(self.seq_maxlen)[None, :] >= n_real[:, None]               # padding at the end
            seq = np.where(pad, self.pad_id if self.pad_id is not None else 0, seq)
            cols.extend(list(seq.T))
        x = np.stack(cols, axis=1).astype(np.int32)
        y = (rng.random_sample((B, 1)) < self.label_rate).astype(np.int16)
        return x, y, domain


def _batch_mixed(self, B, seed, sort=True):
    """(x int32 [B, n_cols], y int16 [B, 1]): rows of MANY domains drawn in proportion to `domain_size`; with `sort`
    the rows are ordered by domain (the domain-sorted, segment-per-domain layout of BASELINE configs[2-3])."""
    x, y, _ = self.batch(B, seed, domain=0)
    rng = np.random.RandomState(seed + 7)
    p = np.asarray(self.domain_size, dtype=np.float64)
    dom = rng.choice(self.n_domain, size=B, p=p / p.sum())
    if sort:
        dom = np.sort(dom, kind="stable")
    x[:, self.domain_idx] = dom.astype(np.int32)
    return x, y


Workload.batch_mixed = _batch_mixed


def amazon_shaped(id_dist="zipf"):
    """BASELINE.json configs[1]: 7 one-hot fields + 2 item-history fields of 5, 25 domains."""
    return Workload("amazon_shaped", AMAZON_FIELD_DIMS, domain_idx=2, itemid_idx=0, n_domain=25,
                    domain_size=AMAZON_DOMAIN_SIZE, n_history_fields=2, seq_maxlen=5, method="mean",
                    pad_id=1368287, label_rate=0.5, id_dist=id_dist)


def aliccp_shaped(id_dist="zipf"):
    """BASELINE.json configs[2]: 23 one-hot fields with the full AliCCP vocabularies, 30 domains."""
    return Workload("aliccp_shaped", ALICCP_FIELD_DIMS, domain_idx=10, itemid_idx=9, n_domain=30,
                    domain_size=ALICCP_DOMAIN_SIZE, label_rate=0.043, id_dist=id_dist)


def cloudtheme_shaped(id_dist="zipf"):
    """BASELINE.json configs[3]: 5 fields, 355 domains.  The dataset is not bundled with the
    reference; vocabulary sizes are assumptions."""
    n_domain = 355
    sizes = (1e6 / np.arange(1, n_domain + 1) ** 1.1).astype(np.int64) + 50
    return Workload("cloudtheme_shaped", [720000, 1360000, n_domain, 1000, 100], domain_idx=2, itemid_idx=1,
                    n_domain=n_domain, domain_size=sizes.tolist(), label_rate=0.05, id_dist=id_dist)


def embedding_stress(total_rows=100_000_000, id_dist="uniform"):
    """BASELINE.json configs[4]: 100 M rows split over 23 fields in the AliCCP ratios."""
    ratio = np.asarray(ALICCP_FIELD_DIMS, dtype=np.float64)
    dims = np.maximum((ratio / ratio.sum() * total_rows).astype(np.int64), 3)
    dims[10] = 30
    return Workload("embedding_stress", dims.tolist(), domain_idx=10, itemid_idx=9, n_domain=30,
                    domain_size=ALICCP_DOMAIN_SIZE, label_rate=0.043, id_dist=id_dist)


WORKLOADS = {"amazon": amazon_shaped, "aliccp": aliccp_shaped, "cloudtheme": cloudtheme_shaped,
             "stress": embedding_stress}
