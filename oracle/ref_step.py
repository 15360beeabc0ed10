"""CPU restatement of the reference's train / eval step.   *** TEST INFRASTRUCTURE ***

What ``make train trainer=cpu`` executes per batch, minus Lightning/Hydra (absent from the
image): models/retrieval.py:80-146 (training_step) and :21-48 + :165-169 (retrieve), built from
``oracle/reference_port.py`` — i.e. the reference's own formulation: Python-loop jagged ops,
padded (B,H,N,N) attention, (B,N,N) bucket tensor per layer, materialised (N',R,D) negatives,
mm + topk.  fp32, torch CPU threads.  Used (a) by parity tests as the end-to-end checker and
(b) by bench.py as the timed CPU baseline (``cpu_baseline`` / ``--impl reference``, kind "port":
the Python reference itself cannot travel to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py may import this file.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import reference_port as O


class RefRetrieval(torch.nn.Module):
    """Parameters use the same names as the B200 pipeline so state dicts are interchangeable."""

    def __init__(self, cfg, all_item_ids: torch.Tensor):
        super().__init__()
        self.cfg = cfg
        D = cfg.embedding_dim
        self.all_item_ids = all_item_ids
        self.params = torch.nn.ParameterDict()
        self.state: Dict[str, torch.Tensor] = {}

    @staticmethod
    def from_state_dict(cfg, all_item_ids, sd: Dict[str, torch.Tensor]) -> "RefRetrieval":
        m = RefRetrieval(cfg, all_item_ids)
        for k, v in sd.items():
            v = v.detach().cpu()
            if v.is_floating_point():
                m.params[k.replace(".", "|")] = torch.nn.Parameter(v.clone().float())
            else:
                m.state[k] = v.clone()
        return m

    def p(self, name: str) -> torch.Tensor:
        return self.params[name.replace(".", "|")]

    def sd(self) -> Dict[str, torch.Tensor]:
        return {k.replace("|", "."): v for k, v in self.params.items()}

    # embeddings.py:94-97
    def item_emb(self, ids: torch.Tensor) -> torch.Tensor:
        e = F.embedding(ids, self.p("embeddings._item_emb.weight"), padding_idx=0)
        if self.cfg.split_year_embedding:
            years = self.state["embeddings.year_lookup_table"][ids.clamp(
                0, self.state["embeddings.year_lookup_table"].numel() - 1)]
            e = torch.cat([e, F.embedding(years, self.p("embeddings._year_emb.weight"),
                                          padding_idx=0)], dim=-1)
        return e

    def features(self, row):
        c = self.cfg
        pad = c.gr_output_length + 1
        lengths = row["history_lengths"]
        ids = F.pad(row["historical_ids"], (0, pad))
        ts = F.pad(row["historical_timestamps"], (0, pad))
        ts = ts.scatter(1, lengths.view(-1, 1), row["target_timestamps"].view(-1, 1))
        return lengths, ids, ts

    # generative_recommenders.py:368-393
    def encode(self, lengths, ids, emb, ts, training: bool):
        c = self.cfg
        D = c.embedding_dim
        N = ids.shape[1]
        x = emb * (D ** 0.5) + self.p("preprocessor._pos_emb.weight")[:N].unsqueeze(0)
        x = F.dropout(x, p=c.dropout, training=training)
        x = x * (ids != 0).unsqueeze(-1).float()
        sd = {k[len("sequence_encoder."):]: v for k, v in self.sd().items()
              if k.startswith("sequence_encoder.")}
        y = O.hstu_forward(lengths, x, ts, sd, c.num_blocks, c.num_heads, c.attention_dim,
                           c.linear_dim, dropout_p=c.dropout, training=training)
        return O.l2_normalize(y, c.l2_eps)

    def training_loss(self, row, neg_draw: Optional[torch.Tensor] = None) -> torch.Tensor:
        """neg_draw: optional pre-drawn sampler offsets (N', R) so a GPU run can be replayed."""
        c = self.cfg
        lengths, ids, ts = self.features(row)
        ids = ids.scatter(1, lengths.view(-1, 1), row["target_ids"].view(-1, 1))
        emb = self.item_emb(ids)
        seq = self.encode(lengths, ids, emb, ts, training=self.training)
        off = O.complete_cumsum(lengths)
        out_emb = O.dense_to_jagged(seq[:, :-1], off)
        sup_ids = O.dense_to_jagged(ids[:, 1:].unsqueeze(-1).float(), off).squeeze(1).long()
        sup_emb = O.dense_to_jagged(emb[:, 1:], off)
        sup_w = O.dense_to_jagged((ids[:, 1:] != 0).float().unsqueeze(-1), off).squeeze(1)
        n = sup_ids.numel()
        R = c.num_negatives
        if c.sampler == "inbatch":
            flat = ids.reshape(-1)
            cid, cemb = O.inbatch_process(flat, flat != 0, self.item_emb(flat), c.l2_eps, True)
            offs = neg_draw if neg_draw is not None else torch.randint(0, cid.numel(), (n, R))
            neg_ids, neg_emb, normed = cid[offs], cemb[offs], True
        else:
            offs = neg_draw if neg_draw is not None else torch.randint(0, self.all_item_ids.numel(), (n, R))
            neg_ids = self.all_item_ids[offs]
            neg_emb, normed = self.item_emb(neg_ids), False
        loss, _ = O.sampled_softmax_loss(out_emb, sup_ids, sup_emb, sup_w, neg_ids, neg_emb,
                                         c.temperature, c.l2_eps, neg_already_normalized=normed)
        return loss

    @torch.no_grad()
    def retrieve(self, row, filter_past_ids: bool = True):
        c = self.cfg
        lengths, ids, ts = self.features(row)
        seq = self.encode(lengths, ids, self.item_emb(ids), ts, training=False)
        cur = O.get_current_embeddings(lengths, seq)
        table = O.l2_normalize(self.item_emb(self.all_item_ids), c.l2_eps)
        return O.candidate_index_topk(cur, table, self.all_item_ids, c.top_k,
                                      ids if filter_past_ids else None)


def torch_topk_baseline(queries: torch.Tensor, items: torch.Tensor, k: int, chunk: int = 1 << 17):
    """The reference's retrieval inner loop as it is (top_k.py:62-69): fp32 mm + torch.topk,
    chunked over the corpus only so that the (B, X) logits fit in host memory."""
    q = queries.float()
    best_s = best_i = None
    for lo in range(0, items.shape[0], chunk):
        s = q @ items[lo:lo + chunk].float().t()
        ts, ti = torch.topk(s, k=min(k, s.shape[1]), dim=1, sorted=True, largest=True)
        ti = ti + lo
        if best_s is None:
            best_s, best_i = ts, ti
        else:
            cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti], 1)
            best_s, sel = torch.topk(cs, k=min(k, cs.shape[1]), dim=1, sorted=True)
            best_i = torch.gather(ci, 1, sel)
    return best_s, best_i
