"""AutoAWQ (GEMM layout) -> optimum format.  Reference: weight_only/utility.py:1245-1459
(`awq_reverse_reorder_int_tensor`, `unpack_awq`, `pack_from_tensors`, `repack_awq_to_optimum_format`), called when an
AutoAWQ checkpoint is loaded (transformers/quantization/utils.py:702).

The reference de-quantises the AWQ tensors to fp16 and re-quantises them (`round((W + z*s) / s)`), a Python loop over the
input channels; the codes come back unchanged, so the whole thing is a nibble permutation, done here with integer tensor
ops on whatever device the checkpoint tensors live on (load-time plumbing, not a hot path):

    AutoAWQ   qweight int32 [K, N/8]   nibble i of word (k, np) = code of column 8*np + ORDER[i], ORDER = [0,2,4,6,1,3,5,7]
              qzeros  int32 [G, N/8]   same packing, zero-points stored as they are
    optimum   qweight int32 [K/8, N]   nibble e of word (kw, n) = code of row 8*kw + e
              qzeros  int32 [G, N/8]   natural nibble order, zero-points stored MINUS ONE (modules.py:363-364)
"""
import torch

# nibble position that holds column c of an AutoAWQ word (the inverse of ORDER)
_POS_OF_COL = (0, 4, 1, 5, 2, 6, 3, 7)


def _unpack_awq_words(packed: torch.Tensor) -> torch.Tensor:
    """int32 [R, N/8] -> uint8-valued int32 [R, N] in natural column order."""
    shifts = torch.tensor([4 * p for p in _POS_OF_COL], dtype=torch.int32, device=packed.device)
    vals = (packed.to(torch.int32).unsqueeze(-1) >> shifts) & 0xF     # [R, N/8, 8], last axis = column within the word
    return vals.reshape(packed.shape[0], packed.shape[1] * 8)


def _pack_along(vals: torch.Tensor, dim: int) -> torch.Tensor:
    """Pack 4-bit values (int32) 8 per word along `dim`, nibble e = element 8*w + e."""
    v = vals.to(torch.int64).movedim(dim, -1)
    v = v.reshape(*v.shape[:-1], v.shape[-1] // 8, 8)
    shifts = torch.arange(0, 32, 4, dtype=torch.int64, device=vals.device)
    words = ((v & 0xF) << shifts).sum(-1)
    words = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
    return words.movedim(-1, dim).contiguous()


def repack_awq_to_optimum_format(awq_qweight: torch.Tensor, awq_qzeros: torch.Tensor, awq_scales: torch.Tensor, bits: int,
                                 group_size: int):
    """utility.py:1432-1459.  Returns (qweight [K/8, N], qzeros [G, N/8], scales [G, N]) in the optimum format."""
    assert bits == 4, "AutoAWQ checkpoints are 4-bit"
    K = awq_qweight.shape[0]
    assert K % 8 == 0 and awq_qzeros.shape[0] == awq_scales.shape[0] and K == awq_scales.shape[0] * group_size
    codes = _unpack_awq_words(awq_qweight)             # [K, N]
    zeros = _unpack_awq_words(awq_qzeros)              # [G, N]
    qweight = _pack_along(codes, 0)                    # [K/8, N]
    qzeros = _pack_along((zeros - 1) & 0xF, 1)         # [G, N/8]
    return qweight, qzeros, awq_scales
