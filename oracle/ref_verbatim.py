"""The reference's own modules on the host CPU, driven without Lightning / Hydra.
*** TEST / BENCH INFRASTRUCTURE ***

Every class and function that does arithmetic here is imported, unmodified, from the staged copy of
the reference package (``oracle/_ref``, written by ``oracle/stage_ref.py``): embeddings, positional
preprocessor, ``HSTU`` with the stock Python fallbacks of the jagged ops (that *is* ``trainer=cpu``),
post-processor, samplers, ``DotProductSimilarity``, ``SampledSoftmaxLoss``, ``CandidateIndex`` +
``MIPSBruteForceTopK``.  What is restated is the glue Lightning would run, line for line:
``Retrieval.training_step`` (models/retrieval.py:80-146), ``GenerativeRecommenders.forward`` and
``.dense_to_jagged`` (models/generative_recommenders.py:368-393, 395-425) — SURVEY.md Appendix A.
``models/retrieval.py`` itself cannot be imported (it imports ``lightning`` and ``hydra``).

Used by ``bench.py --impl reference`` and its ``cpu_baseline`` leg (kind "reference"), and by
tests/test_oracle_golden.py to check the port against it.  Nothing in the product imports this.
"""
from __future__ import annotations

import logging
import sys
from pathlib import Path

import torch

REF_ROOT = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_ROOT / "generative_recommenders_pl" / "models" / "sequential_encoders" / "hstu.py").exists()


def _import():
    if not available():
        raise RuntimeError("oracle/_ref is not staged (python -m oracle.stage_ref where /root/reference exists)")
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    logging.disable(logging.CRITICAL)          # the reference logs every fbgemm fallback
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):   # embeddings.py prints a missing-CSV warning
        from generative_recommenders_pl.models.embeddings.embeddings import LocalEmbeddingModule
    from generative_recommenders_pl.models.indexing.candidate_index import CandidateIndex
    from generative_recommenders_pl.models.indexing.top_k import MIPSBruteForceTopK
    from generative_recommenders_pl.models.losses.autoregressive_losses import SampledSoftmaxLoss
    from generative_recommenders_pl.models.negatives_samples.negative_sampler import (
        InBatchNegativesSampler, LocalNegativesSampler)
    from generative_recommenders_pl.models.postprocessors.postprocessors import L2NormEmbeddingPostprocessor
    from generative_recommenders_pl.models.preprocessors import (
        LearnablePositionalEmbeddingInputFeaturesPreprocessor)
    from generative_recommenders_pl.models.sequential_encoders.hstu import HSTU
    from generative_recommenders_pl.models.similarity.dot_product import DotProductSimilarity
    from generative_recommenders_pl.models.utils import ops
    from generative_recommenders_pl.models.utils.features import seq_features_from_row
    return dict(locals())


class ReferenceRetrieval(torch.nn.Module):
    """``Retrieval`` (models/retrieval.py) assembled from the reference's classes with the kwargs
    ``GenerativeRecommenders.__init__`` injects (generative_recommenders.py:118-239)."""

    def __init__(self, cfg, all_item_ids: torch.Tensor):
        super().__init__()
        R = _import()
        self._ops, self._features = R["ops"], R["seq_features_from_row"]
        self.cfg = cfg
        D, N = cfg.embedding_dim, cfg.N
        self.embeddings = R["LocalEmbeddingModule"](num_items=cfg.num_items, item_embedding_dim=D)
        self.preprocessor = R["LearnablePositionalEmbeddingInputFeaturesPreprocessor"](
            max_sequence_len=N, embedding_dim=D, dropout_rate=cfg.dropout)
        self.sequence_encoder = R["HSTU"](
            max_sequence_len=cfg.max_sequence_length, max_output_len=cfg.gr_output_length + 1,
            embedding_dim=D, item_embedding_dim=D, num_blocks=cfg.num_blocks, num_heads=cfg.num_heads,
            linear_dim=cfg.linear_dim, attention_dim=cfg.attention_dim, normalization="rel_bias",
            linear_config="uvqk", linear_activation="silu", linear_dropout_rate=cfg.dropout,
            attn_dropout_rate=0.0)
        self.postprocessor = R["L2NormEmbeddingPostprocessor"](embedding_dim=D, eps=cfg.l2_eps)
        if cfg.sampler == "inbatch":
            self.negatives_sampler = R["InBatchNegativesSampler"](
                l2_norm=True, l2_norm_eps=cfg.l2_eps, dedup_embeddings=True)
        else:
            self.negatives_sampler = R["LocalNegativesSampler"](
                l2_norm=True, l2_norm_eps=cfg.l2_eps, all_item_ids=all_item_ids.tolist())
        self._inbatch = cfg.sampler == "inbatch"
        self.similarity = R["DotProductSimilarity"]()
        self.loss = R["SampledSoftmaxLoss"](num_to_sample=cfg.num_negatives,
                                            softmax_temperature=cfg.temperature)
        self.candidate_index = R["CandidateIndex"](k=cfg.top_k, ids=all_item_ids,
                                                   top_k_module=R["MIPSBruteForceTopK"]())

    @staticmethod
    def complete_row(row):
        """``seq_features_from_row`` also reads ratings / years; the synthetic generator has none."""
        row = dict(row)
        for k, like in (("historical_ratings", "historical_ids"), ("historical_years", "historical_ids"),
                        ("target_ratings", "target_ids"), ("target_years", "target_ids")):
            row.setdefault(k, torch.zeros_like(row[like]))
        return row

    def forward(self, sf):                                   # generative_recommenders.py:368-393
        pl, ue, vm, _ = self.preprocessor(past_lengths=sf.past_lengths, past_ids=sf.past_ids,
                                          past_embeddings=sf.past_embeddings, past_payloads=sf.past_payloads)
        ue, cache = self.sequence_encoder(past_lengths=pl, user_embeddings=ue, valid_mask=vm,
                                          past_payloads=sf.past_payloads)
        return self.postprocessor(ue), cache

    def dense_to_jagged(self, lengths, **kwargs):            # generative_recommenders.py:395-425
        ops = self._ops
        off = ops.asynchronous_complete_cumsum(lengths)
        out = {}
        if "supervision_ids" in kwargs:
            out["supervision_ids"] = ops.dense_to_jagged(
                kwargs.pop("supervision_ids").unsqueeze(-1).float(), off).squeeze(1).long()
        if "supervision_weights" in kwargs:
            out["supervision_weights"] = ops.dense_to_jagged(
                kwargs.pop("supervision_weights").unsqueeze(-1), off).squeeze(1)
        for key, value in kwargs.items():
            out[key] = ops.dense_to_jagged(value, off)
        return out

    def training_loss(self, row) -> torch.Tensor:            # retrieval.py:80-133
        dev = next(self.parameters()).device
        sf, target_ids, _ = self._features(self.complete_row(row), device=dev,
                                           max_output_length=self.cfg.gr_output_length + 1)
        sf.past_ids.scatter_(dim=1, index=sf.past_lengths.view(-1, 1), src=target_ids.view(-1, 1))
        input_embeddings = self.embeddings.get_item_embeddings(sf.past_ids)
        sf = sf._replace(past_embeddings=input_embeddings)
        seq_embeddings, _ = self.forward(sf)
        supervision_ids = sf.past_ids
        if self._inbatch:
            in_batch_ids = supervision_ids.view(-1)
            self.negatives_sampler.process_batch(
                ids=in_batch_ids, presences=(in_batch_ids != 0),
                embeddings=self.embeddings.get_item_embeddings(in_batch_ids))
        else:
            self.negatives_sampler._embeddings_module = self.embeddings
        jf = self.dense_to_jagged(
            lengths=sf.past_lengths, output_embeddings=seq_embeddings[:, :-1, :],
            supervision_ids=supervision_ids[:, 1:], supervision_embeddings=input_embeddings[:, 1:, :],
            supervision_weights=(supervision_ids[:, 1:] != 0).float())
        return self.loss.jagged_forward(negatives_sampler=self.negatives_sampler,
                                        similarity=self.similarity, **jf)


def reference_topk(queries: torch.Tensor, items: torch.Tensor, k: int, chunk_queries: int = 256):
    """``MIPSBruteForceTopK.forward`` (models/indexing/top_k.py:44-70) itself: mm + topk + id gather
    over the whole corpus, called on slices of the query batch so that the (B, X) logits fit in
    host memory."""
    R = _import()
    mod = R["MIPSBruteForceTopK"]()
    items_t = items.float().t()
    ids = torch.arange(1, items.shape[0] + 1).unsqueeze(0)
    out = []
    for lo in range(0, queries.shape[0], chunk_queries):
        out.append(mod(query_embeddings=queries[lo:lo + chunk_queries].float(), item_embeddings_t=items_t,
                       item_ids=ids, k=k, sorted=True))
    return torch.cat([o[0] for o in out]), torch.cat([o[1] for o in out])
