"""Autoregressive losses — drop-in for ``SampledSoftmaxLoss`` of the reference's
``models/losses/autoregressive_losses.py`` (:249-306).

``jagged_forward`` keeps the reference signature.  When the sampler can hand out table row
indices (``fused_sample``) and the similarity is a plain dot product, the whole block
  sample -> gather (N',R,D) -> L2 norm -> bmm -> /T -> collision mask -> log_softmax -> mean
runs as ONE forward and ONE backward kernel (``grb_sampled_softmax_fwd/bwd``).  Any other
(sampler, similarity) pair takes the unfused path, which is the reference's math on torch ops.
"""
from __future__ import annotations

import abc

import torch
import torch.nn.functional as F

from . import functional as GF
from .negative_sampler import NegativesSampler
from .similarity import DotProductSimilarity, NDPModule


class AutoregressiveLoss(torch.nn.Module):
    @abc.abstractmethod
    def jagged_forward(self, output_embeddings: torch.Tensor, supervision_ids: torch.Tensor,
                       supervision_embeddings: torch.Tensor, supervision_weights: torch.Tensor,
                       negatives_sampler: NegativesSampler) -> torch.Tensor:
        pass


class _ZeroRows(torch.autograd.Function):
    """Identity whose backward zeroes one row of the gradient (nn.Embedding padding_idx)."""

    @staticmethod
    def forward(ctx, table, row):
        ctx.row = row
        return table.view_as(table)

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        g[ctx.row].zero_()
        return g, None


def _is_plain_dot(similarity) -> bool:
    if isinstance(similarity, DotProductSimilarity):
        return True
    dbg = getattr(similarity, "debug_str", None)
    return type(similarity).__name__ == "DotProductSimilarity" and callable(dbg) and dbg() == "dp"


class SampledSoftmaxLoss(AutoregressiveLoss):
    def __init__(self, num_to_sample: int, softmax_temperature: float, bf16_backward: bool = False) -> None:
        """``bf16_backward`` (extension, default off = the reference's fp32 arithmetic): with the in-batch
        sampler the backward of the fused loss reads bf16 copies of the output embeddings and of the
        cache rows (GF.sampled_softmax_rows) — for encoders that compute in bf16 anyway."""
        super().__init__()
        self._num_to_sample: int = num_to_sample
        self._softmax_temperature: float = softmax_temperature
        self._bf16_backward: bool = bool(bf16_backward)

    def jagged_forward(self, output_embeddings: torch.Tensor, supervision_ids: torch.Tensor,
                       supervision_embeddings: torch.Tensor, supervision_weights: torch.Tensor,
                       negatives_sampler: NegativesSampler, similarity: NDPModule) -> torch.Tensor:
        assert output_embeddings.size() == supervision_embeddings.size()
        assert supervision_ids.size() == supervision_embeddings.size()[:-1]
        assert supervision_ids.size() == supervision_weights.size()

        fused = None
        if (_is_plain_dot(similarity) and output_embeddings.is_cuda
                and output_embeddings.dtype == torch.float32 and output_embeddings.dim() == 2
                and (output_embeddings.size(1) <= 256 or output_embeddings.size(1) == 512)
                and supervision_ids.size(0) > 1):
                # a reference sampler (no fused_sample) paired with this loss takes the composite below
            fused_sample = getattr(negatives_sampler, "fused_sample", None)
            fused = fused_sample(supervision_ids, self._num_to_sample) if fused_sample is not None else None
            if fused is not None and fused.table0.dtype != torch.float32:
                raise NotImplementedError("fused sampled softmax needs float32 tables")
        positive_embeddings = negatives_sampler.normalize_embeddings(supervision_embeddings)
        if fused is not None:
            t0, t1 = fused.table0, fused.table1
            if fused.zero_grad_rows[0] is not None:
                t0 = _ZeroRows.apply(t0, fused.zero_grad_rows[0])
            if t1 is not None and fused.zero_grad_rows[1] is not None:
                t1 = _ZeroRows.apply(t1, fused.zero_grad_rows[1])
            jagged_loss = GF.sampled_softmax_rows(
                output_embeddings, positive_embeddings, t0, t1, fused.idx0, fused.idx1,
                supervision_ids, fused.ids, fused.l2_norm, negatives_sampler._l2_norm_eps,
                self._softmax_temperature, bf16_backward=self._bf16_backward)
            return GF.weighted_mean(jagged_loss, supervision_weights)

        # unfused: the reference's sequence on torch ops (autoregressive_losses.py:272-306)
        sampled_ids, sampled_negative_embeddings = negatives_sampler(
            positive_ids=supervision_ids, num_to_sample=self._num_to_sample)
        positive_logits = similarity(
            input_embeddings=output_embeddings, item_embeddings=positive_embeddings.unsqueeze(1),
            item_sideinfo=None, item_ids=supervision_ids.unsqueeze(1),
            precomputed_logits=None) / self._softmax_temperature
        negative_logits = similarity(
            input_embeddings=output_embeddings, item_embeddings=sampled_negative_embeddings,
            item_sideinfo=None, item_ids=sampled_ids, precomputed_logits=None)
        negative_logits = torch.where(supervision_ids.unsqueeze(1) == sampled_ids, -5e4,
                                      negative_logits / self._softmax_temperature)
        jagged_loss = -F.log_softmax(torch.cat([positive_logits, negative_logits], dim=1), dim=1)[:, 0]
        return (jagged_loss * supervision_weights).sum() / supervision_weights.sum()
