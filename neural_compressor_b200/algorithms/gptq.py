"""GPTQ quantizer.  Reference: neural_compressor/torch/algorithms/weight_only/gptq.py
    trace_gptq_target_blocks :68-107, find_layers :109-131
    RAWGPTQuantizer.prepare_for_calibration :398-458, execute_quantization :567-1086
    GPTQ.add_batch :1111-1141, fasterquant :1143-1351, Quantizer.find_params :1501-1624
    GPTQuantizer (the `Quantizer` subclass the algorithm entry drives) :1651-1716

Host Python keeps the reference's orchestration (capture block-0 inputs by raising inside the patched
forward, block by block: hooks -> Hessians -> fasterquant -> re-run the block with quantised weights ->
export).  Everything numeric runs on the B200:
    K1  b200woq_hessian_accumulate   one accumulator per DISTINCT input tensor (q/k/v and gate/up share theirs;
                                     the reference recomputes the identical X^T X for each, gptq.py:665-678)
    K2  Cholesky inverse factor      cuSOLVER through torch.linalg on the device (round 1)
    K3  b200woq_gptq_fasterquant     column loop + lazy updates, emits codes / scales / zeros / fake-quant
    K4  b200woq_pack_codes/_params   optimum-format packing on the device (the reference packs on the CPU, :838)
"""
from __future__ import annotations

import os
import contextlib
import time
from functools import partial
from typing import Dict, List

import torch

from .. import ops
from ..utils import current_device, find_layers, get_model_device, logger, move_to_device, set_module
from .base_algorithm import Quantizer
from .modules import B200WeightOnlyLinear


def _is_conv1d(m) -> bool:
    try:
        import transformers

        return isinstance(m, transformers.Conv1D)
    except Exception:  # pragma: no cover
        return False


def trace_gptq_target_blocks(model):
    """gptq.py:68-107: the first nn.ModuleList is the transformer stack; what precedes it are embeddings."""
    info = {"embeddings": {}, "transformers_pre": {}, "transformers_name": "", "transformers": [], "transformers_post": {}}
    found = False
    for name, module in model.named_modules():
        if not found and type(module) is torch.nn.ModuleList:
            info["transformers_name"] = name
            info["transformers"] = module
            found = True
        elif not found and len(list(module.children())) == 0:
            info["embeddings"][name] = module
        elif found and name.find(info["transformers_name"]) == -1:
            # gptq.py:122-125: every module after the stack overwrites the entry, so the LAST one (lm_head) wins
            info["transformers_post"] = {"name": name, "layer": module}
    if not found:
        raise ValueError("GPTQ needs a model with an nn.ModuleList of transformer blocks")
    return info


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of the calibration sequences owned by `rank` (strong scaling over samples)."""
    return rank * n_items // world, (rank + 1) * n_items // world


class _HessianBank:
    """One raw accumulator (sum of X^T X) per distinct input tensor of a block forward."""

    def __init__(self, device):
        self.device = device
        self.acc: List[torch.Tensor] = []
        self.layer_to_slot: Dict[str, int] = {}
        self.nsamples = 0
        self._seen = {}  # input identity -> (slot, tensor), valid during one block forward
        self._fwd_batch = 0

    def begin_forward(self):
        self._seen = {}
        self._fwd_batch = 0

    def end_forward(self):
        # GPTQ.add_batch counts the batch dimension of the layer input (gptq.py:1118, 1136-1137)
        self.nsamples += max(self._fwd_batch, 1)

    def add(self, layer_name: str, inp: torch.Tensor):
        # identity of the input tensor: (address, shape, dtype).  The tensor is kept alive until end_forward()
        # so the caching allocator cannot hand the same address to a later activation of this forward.
        key = (inp.data_ptr(), tuple(inp.shape), inp.dtype)
        self._fwd_batch = max(self._fwd_batch, inp.shape[0] if inp.dim() == 3 else 1)
        slot = self.layer_to_slot.get(layer_name)
        if key in self._seen:
            shared = self._seen[key][0]
            if slot is None:
                self.layer_to_slot[layer_name] = shared
            elif slot != shared:
                raise RuntimeError(f"inconsistent input sharing for {layer_name}")
            return
        if slot is None:
            c = inp.shape[-1]
            self.acc.append(torch.zeros((c, c), dtype=torch.float32, device=self.device))
            slot = len(self.acc) - 1
            self.layer_to_slot[layer_name] = slot
        elif any(v[0] == slot for v in self._seen.values()):
            raise RuntimeError(f"inconsistent input sharing for {layer_name}")
        self._seen[key] = (slot, inp)
        x = inp if inp.is_contiguous() else inp.contiguous()
        ops.hessian_accumulate(x, self.acc[slot])

    def all_reduce(self):
        """Calibration data-parallelism: H is additive over samples, so each rank accumulates its
        shard of the sequences and the raw sums are all-reduced once per distinct input (NCCL over NVLink)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        for h in self.acc:
            dist.all_reduce(h, op=dist.ReduceOp.SUM)
        n = torch.tensor([float(self.nsamples)], device=self.acc[0].device if self.acc else "cpu")
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        self.nsamples = int(round(n.item()))

    def reduce_to_owners(self, owner_of_slot: Dict[int, int]):
        """Each raw accumulator is summed onto ONE owner rank (`dist.reduce`), which alone factorises it and then
        broadcasts the inverse factor: the Cholesky chain is no longer replicated on every rank and there is no host
        synchronisation here (the global sample count is cached by the engine after the first block)."""
        import torch.distributed as dist

        for slot, h in enumerate(self.acc):
            dist.reduce(h, dst=owner_of_slot[slot], op=dist.ReduceOp.SUM)


def slot_owners(n_slots: int, world: int) -> Dict[int, int]:
    """Owner rank of each distinct Hessian of a block, in order of first use (Llama: q/k/v, o, gate/up, down)."""
    return {i: i % world for i in range(n_slots)}


def _dist_world():
    import torch.distributed as dist

    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _dist_rank():
    import torch.distributed as dist

    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


class RAWGPTQuantizer:
    """The engine behind GPTQuantizer (gptq.py:184-1086), default flags of the reference honoured."""

    def __init__(self, model, weight_config=None, nsamples=128, use_max_length=True, max_seq_length=2048, device=None,
                 use_layer_wise=False, model_path="", quant_lm_head=False, use_block_wise=False, **kwargs):
        self.model = model
        self.blocks_info = trace_gptq_target_blocks(model)
        self.dtype = next(iter(model.parameters())).dtype
        self.weight_config = weight_config or {}
        self.device = current_device()
        self.quant_lm_head = quant_lm_head
        if use_layer_wise or use_block_wise:
            logger.info("use_layer_wise / use_block_wise are host-RAM saving modes of the reference; ignored on the B200")
        self._check_layer_config()
        self.timing = {"hessian_fwd": 0.0, "cholesky": 0.0, "fasterquant": 0.0, "propagate": 0.0, "pack": 0.0}
        self.profile = False
        self.offload_packed_to_host = False  # reference behaviour `transformer_block.cpu()` (gptq.py:766)

    def _check_layer_config(self):
        """gptq.py:337-365 defaults."""
        for cfg in self.weight_config.values():
            cfg.setdefault("dtype", "int")
            cfg.setdefault("bits", 4)
            cfg.setdefault("group_size", 128)
            cfg.setdefault("block_size", cfg["group_size"])
            cfg.setdefault("percdamp", 0.01)
            cfg.setdefault("sym", False)
            cfg.setdefault("act_order", False)
            cfg.setdefault("static_groups", False)
            cfg.setdefault("true_sequential", False)
            cfg.setdefault("mse", False)
            if cfg["dtype"] != "int" and "int" in cfg["dtype"]:
                cfg["bits"] = int(cfg["dtype"].lstrip("int"))
                cfg["dtype"] = "int"

    def get_layer_config(self, layer_name):
        """gptq.py:367-382: exact name first; otherwise the FIRST entry (the reference's regex test
        `len(findall) is not None` is always true -- SURVEY §8 a18)."""
        cfg = self.weight_config.get(layer_name)
        if cfg is not None:
            return cfg
        for _, v in self.weight_config.items():
            return v
        return None

    def full_name(self, sub_layer_name, block_idx):
        return ".".join([self.blocks_info["transformers_name"], str(block_idx), sub_layer_name])

    # ------------------------------------------------------------------ calibration capture
    @torch.no_grad()
    def prepare_for_calibration(self):
        """gptq.py:398-458: record (*args, **kwargs) of block 0 per sample, abort the forward with ValueError."""
        self.cache_kwargs = {"batch_num": 0}
        self.cache_args: List[list] = []
        blocks = self.blocks_info["transformers"]

        def capture(layer, *args, **kwargs):
            self.cache_kwargs["batch_num"] += 1
            for k, v in kwargs.items():
                if isinstance(v, torch.Tensor) or k in ("alibi", "position_embeddings"):
                    self.cache_kwargs.setdefault(k, []).append(v)
            for idx, item in enumerate(args):
                if idx + 1 > len(self.cache_args):
                    self.cache_args.append([])
                self.cache_args[idx].append(item)
            raise ValueError

        for emb in self.blocks_info["embeddings"].values():
            emb.to(self.device)
        # (a layer-sharded model may hold block 0 as a storage-less skeleton on this rank: only its forward is
        # intercepted here, its weights arrive with the owner's broadcast in quantize_block)
        if not any(p.is_meta for p in blocks[0].parameters()):
            blocks[0] = blocks[0].to(self.device)
        self._block0_forward = blocks[0].forward
        blocks[0].forward = partial(capture, blocks[0])
        self._model_forward = self.model.forward
        orig = self.model.forward
        dev = self.device

        def model_forward(model, *args, **kwargs):
            try:
                orig(*move_to_device(args, dev), **move_to_device(kwargs, dev))
            except ValueError:
                pass

        self.model.forward = partial(model_forward, self.model)

    @torch.no_grad()
    def remove_prepare_for_calibration(self):
        self.model.forward = self._model_forward
        self.blocks_info["transformers"][0].forward = self._block0_forward

    @staticmethod
    def _same(a, b) -> bool:
        if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
            return a.shape == b.shape and a.dtype == b.dtype and (a.data_ptr() == b.data_ptr() or torch.equal(a, b))
        if isinstance(a, (tuple, list)) and isinstance(b, (tuple, list)) and len(a) == len(b):
            return all(RAWGPTQuantizer._same(x, y) for x, y in zip(a, b))
        return a is b or a == b

    def rechunk_calibration(self, batch=None):
        """Process `batch` cached sequences per block forward instead of one (the reference runs them one by one,
        gptq.py:680-685).  Sequences are independent inside a decoder block, so the Hessians and the propagated
        outputs are the same up to GEMM rounding; the forward GEMMs and the Hessian launches get 8x larger tiles.
        Only done when every sequence has the same shape and identical auxiliary inputs (position embeddings, mask);
        B200WOQ_CALIB_BATCH=1 restores the one-by-one schedule."""
        if getattr(self, "_chunked", False):
            return
        self._chunked = True
        B = int(os.environ.get("B200WOQ_CALIB_BATCH", "8")) if batch is None else batch
        n = self.cache_kwargs.get("batch_num", 0)
        if B <= 1 or n <= 1:
            return
        in_kwargs = "hidden_states" in self.cache_kwargs
        hs = self.cache_kwargs["hidden_states"] if in_kwargs else (self.cache_args[0] if self.cache_args else None)
        if hs is None or any((h.dim() != 3 or h.shape[0] != 1 or h.shape != hs[0].shape) for h in hs):
            return
        for k, v in self.cache_kwargs.items():
            if k in ("batch_num", "hidden_states"):
                continue
            if len(v) != n or not all(self._same(v[0], x) for x in v[1:]):
                return
        for idx, lst in enumerate(self.cache_args):
            if idx == 0 and not in_kwargs:
                continue
            if len(lst) != n or not all(self._same(lst[0], x) for x in lst[1:]):
                return
        chunks = [torch.cat(hs[i:i + B], dim=0) for i in range(0, n, B)]
        m = len(chunks)
        for k in list(self.cache_kwargs):
            if k not in ("batch_num", "hidden_states"):
                self.cache_kwargs[k] = [self.cache_kwargs[k][0]] * m
        for idx in range(len(self.cache_args)):
            if not (idx == 0 and not in_kwargs):
                self.cache_args[idx] = [self.cache_args[idx][0]] * m
        if in_kwargs:
            self.cache_kwargs["hidden_states"] = chunks
        else:
            self.cache_args[0] = chunks
        self.cache_kwargs["batch_num"] = m
        logger.info(f"calibration forwards batched: {n} sequences -> {m} chunks of <= {B}")

    def _batch(self, j):
        kw = {k: v[j] for k, v in self.cache_kwargs.items()}
        args = [a[j] for a in self.cache_args]
        return args, kw

    @staticmethod
    def _hidden(out):
        return out if isinstance(out, torch.Tensor) else out[0]

    def _sequentials(self, block):
        """gptq.py:538-565 + :631-635."""
        layers = list(find_layers(block))
        true_seq = False
        for cfg in self.weight_config.values():
            if cfg.get("true_sequential") is not None:
                true_seq = cfg["true_sequential"]
                break
        if not true_seq:
            return [layers]
        if "q" in layers[0].lower() and "k" in layers[0].lower():
            qkv, post = [layers[0]], layers[1:]
        else:
            qkv, post = layers[0:3], layers[3:]
        return [qkv] + [[l] for l in post]

    def _fasterquant_rows_sharded(self, W, Hinv, dead, cfg):
        """The rows of a layer are independent given Hinv (per-row scale/zero gptq.py:1544-1565, per-row updates
        :1297-1304), so with several ranks each one runs the column loop on N/world rows.  Only the u8
        codes, the per-group scale / zero and the per-row losses are all-gathered (N*C + 8*N*G + 4*N bytes instead of
        the 5*N*C bytes that shipping the fp32 fake-quant weights would cost); every rank rebuilds
        Q = scale * (code - zero) itself, which is the very expression the column loop stores (gptq.py:1636-1637).
        Bit-identical to the un-sharded result."""
        import torch.distributed as dist

        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        N = W.shape[0]
        dq = self._double_quant(cfg)
        # double quant groups the scales over consecutive output rows: shards must not cut a group
        if world == 1 or N % world != 0 or (dq is not None and (N // world) % dq["group_size"] != 0):
            return ops.gptq_fasterquant(W, Hinv, dead, cfg["block_size"], cfg["group_size"], cfg["bits"], cfg["sym"],
                                        cfg["mse"], double_quant=self._double_quant(cfg))
        rank = dist.get_rank()
        rows = N // world
        part = ops.gptq_fasterquant(W[rank * rows:(rank + 1) * rows].contiguous(), Hinv, dead, cfg["block_size"],
                                    cfg["group_size"], cfg["bits"], cfg["sym"], cfg["mse"], want_q=False, double_quant=dq)
        out = {}
        for k in ("codes", "scale", "zero", "losses"):
            v = part[k]
            full = torch.empty((N,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            dist.all_gather_into_tensor(full, v.contiguous())
            out[k] = full
        out["Q"] = ops.gptq_rebuild_q(out["codes"], out["scale"], out["zero"], cfg["group_size"])
        return out

    @staticmethod
    def _double_quant(cfg):
        if not cfg.get("use_double_quant"):
            return None
        return dict(bits=int(cfg.get("double_quant_bits", 8)), group_size=int(cfg.get("double_quant_group_size", 256)),
                    sym=bool(cfg.get("double_quant_sym", False)))

    def _global_nsamples(self, local_n: int) -> int:
        """Sum of the ranks' sample counts.  Every block sees the same calibration set, so the (synchronising) all-reduce
        runs once, in the first block, and the result is cached."""
        if _dist_world() == 1:
            return local_n
        cached = getattr(self, "_nsamples_cache", None)
        if cached is None or cached[0] != local_n:
            import torch.distributed as dist

            n = torch.tensor([float(local_n)], device=self.device)
            dist.all_reduce(n, op=dist.ReduceOp.SUM)
            cached = self._nsamples_cache = (local_n, int(round(n.item())))
        return cached[1]

    def _share_factor(self, ent, owner: int, C: int, act_order: bool):
        """Owner rank -> everyone: the inverse factor U (fp32 C x C), the dead-column mask, the factorisation status and,
        with act_order, the permutation.  Runs on the current (side) stream; the owner first waits for its K2 chain."""
        import torch.distributed as dist

        cur = torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
        if _dist_rank() == owner:
            if cur is not None and ent["done"] is not None:
                cur.wait_event(ent["done"])
            ent["Hinv"] = ent["Hinv"].contiguous()   # a collective ships storage order (no-op for the kernel's output)
        else:
            ent["Hinv"] = torch.empty((C, C), dtype=torch.float32, device=self.device)
            ent["dead"] = torch.empty(C, dtype=torch.uint8, device=self.device)
            ent["info"] = torch.empty(1, dtype=torch.int32, device=self.device)
            ent["perm"] = torch.empty(C, dtype=torch.int64, device=self.device) if act_order else None
        dist.broadcast(ent["Hinv"], src=owner)
        dist.broadcast(ent["dead"], src=owner)
        dist.broadcast(ent["info"], src=owner)
        if act_order:
            dist.broadcast(ent["perm"], src=owner)
        if cur is not None:
            ent["done"] = torch.cuda.Event()
            ent["done"].record()
        ent["shared"] = True

    def _queue_factor_status(self, infos, block_idx):
        self._check_factor_status()  # at most one status in flight (true_sequential queues several per block)
        st = torch.stack(infos).max().reshape(1)
        if st.is_cuda:
            host = torch.empty(1, dtype=st.dtype, pin_memory=True)
            host.copy_(st, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host, ev = st.clone(), None
        self._pending_status = (ev, host, block_idx)

    def _check_factor_status(self):
        pend = getattr(self, "_pending_status", None)
        if pend is None:
            return
        self._pending_status = None
        ev, host, block_idx = pend
        if ev is not None:
            ev.synchronize()
        if int(host.item()) != 0:
            raise torch.linalg.LinAlgError(
                f"block {block_idx}: a layer-input Hessian is not positive-definite after damping (cholinv status "
                f"{int(host.item())}); raise percdamp")

    def _sync_time(self, key, t0):
        if self.profile:
            torch.cuda.synchronize()
            self.timing[key] += time.perf_counter() - t0
            return time.perf_counter()
        return t0

    # ------------------------------------------------------------------ main loop
    @torch.no_grad()
    def execute_quantization(self):
        """gptq.py:567-1086."""
        blocks = self.blocks_info["transformers"]
        for p in self.model.parameters():
            p.requires_grad = False
        cb = getattr(self, "block_callback", None)  # optional progress hook: called with the index of each finished block
        for block_idx in range(len(blocks)):
            self.quantize_block(block_idx)
            if cb is not None:
                cb(block_idx)
        self._check_factor_status()
        if self.quant_lm_head:
            self.quantize_post_layer()
        return self.model

    @torch.no_grad()
    def quantize_post_layer(self):
        """gptq.py:886-1078 (`quant_lm_head=True`): the module after the transformer stack (lm_head) goes through the same
        Hessian -> factor -> column loop -> pack chain.  As in the reference its calibration inputs are the outputs of the
        LAST block as they sit in the cache -- the final norm of the model is not applied (gptq.py:929-933)."""
        import torch.distributed as dist

        post = self.blocks_info.get("transformers_post") or {}
        name, layer = post.get("name"), post.get("layer")
        if layer is None or not (isinstance(layer, torch.nn.Linear) or _is_conv1d(layer)):
            logger.warning("quant_lm_head: no Linear found after the transformer stack; nothing to do")
            return
        cfg = self.get_layer_config(name)
        if cfg is None:
            logger.warning(f"{name} can be quantized but excluded from quantization configs.")
            return
        if cfg.get("fp8_aware") or cfg.get("static_groups"):
            raise NotImplementedError("fp8_aware / static_groups are not available for the post-transformer layer")
        logger.info("Quantizing post transformer layers")
        layer.to(self.device)
        hs = self.cache_kwargs["hidden_states"] if "hidden_states" in self.cache_kwargs else self.cache_args[0]
        C = hs[0].shape[-1]
        Hsum = torch.zeros((C, C), dtype=torch.float32, device=self.device)
        local_n = 0
        for x in hs:
            x = x.to(self.device)
            ops.hessian_accumulate(x if x.is_contiguous() else x.contiguous(), Hsum)
            local_n += x.shape[0] if x.dim() == 3 else 1
        if _dist_world() > 1:
            dist.all_reduce(Hsum, op=dist.ReduceOp.SUM)
        nsamples = self._global_nsamples(local_n)
        Hc, dead = ops.hessian_finalize(Hsum, nsamples, cfg["percdamp"])
        perm = None
        if cfg["act_order"]:
            perm = torch.argsort(torch.diag(Hc), descending=True)
        elif cfg.get("hybrid_order"):
            perm = hybrid_order_perm(torch.diag(Hc), int(cfg["group_size"]))
        if perm is not None:
            Hc = Hc[perm][:, perm].contiguous()
            dead = dead[perm].contiguous()
        Hinv = ops.cholesky_inverse_upper(Hc)
        W = layer.weight.data
        W = (W.t() if _is_conv1d(layer) else W).float()
        W = (W[:, perm] if perm is not None else W).contiguous().clone()
        r = self._fasterquant_rows_sharded(W, Hinv, dead, cfg)
        g_idx = None
        if perm is not None:
            inv = torch.argsort(perm)
            r["codes"] = r["codes"][:, inv].contiguous()
            if cfg["act_order"]:
                g_idx = perm
            else:
                g = int(cfg["group_size"])
                inv_order = torch.argsort(perm[::g] // g)
                r["scale"], r["zero"] = r["scale"][:, inv_order].contiguous(), r["zero"][:, inv_order].contiguous()
        if _is_conv1d(layer):
            in_f, out_f = layer.weight.shape[0], layer.weight.shape[1]
        else:
            in_f, out_f = layer.in_features, layer.out_features
        new_module = B200WeightOnlyLinear(in_f, out_f, dtype=cfg["dtype"], bits=cfg["bits"], group_size=cfg["group_size"],
                                          zp=not cfg["sym"], bias=layer.bias is not None, g_idx=g_idx is not None,
                                          device=self.device)
        new_module.pack_stored(r["codes"], r["scale"], None if cfg["sym"] else r["zero"], layer.bias, g_idx=g_idx)
        set_module(self.model, name, new_module)

    @torch.no_grad()
    def _side_streams(self, n):
        pool = getattr(self, "_streams", None)
        if pool is None or len(pool) != n:
            pool = self._streams = [torch.cuda.Stream(device=self.device) for _ in range(n)]
        return pool

    def quantize_block(self, block_idx: int):
        blocks = self.blocks_info["transformers"]
        self.rechunk_calibration()
        self._check_factor_status()
        # layer-sharded models (utils/sharded.py, BASELINE configs[4]): the block lives on its owner rank only; the owner
        # broadcasts its weights over NCCL right before the block is processed and the other ranks drop them afterwards
        shard = getattr(self.model, "_b200_shard", None)
        fetched = False
        if shard is not None and shard["world"] > 1:
            from ..utils import sharded

            owner = sharded.block_owner(block_idx, shard["world"], shard["n_blocks"])
            t_f = time.perf_counter()
            nbytes = sharded.fetch_block(blocks[block_idx], owner, self.device)
            fetched = owner != shard["rank"]
            self.fetch_stats = getattr(self, "fetch_stats", [])
            if self.profile:
                torch.cuda.synchronize()
            self.fetch_stats.append((block_idx, nbytes, time.perf_counter() - t_f))
        block = blocks[block_idx].to(self.device)
        sub_layers = find_layers(block)
        for sequential in self._sequentials(block):
            layers = {n: sub_layers[n] for n in sequential
                      if self.get_layer_config(self.full_name(n, block_idx)) is not None}
            # ---- forward pass #1: Hessians (gptq.py:670-687) ----
            t0 = time.perf_counter()
            bank = _HessianBank(self.device)
            handles = []
            for lname, layer in layers.items():
                handles.append(layer.register_forward_hook(
                    lambda _m, inp, _out, _n=lname: bank.add(_n, inp[0].data)))
            batch_num = self.cache_kwargs.pop("batch_num")
            for j in range(batch_num):
                args, kw = self._batch(j)
                bank.begin_forward()
                block(*args, **kw)
                bank.end_forward()
            self.cache_kwargs["batch_num"] = batch_num
            for h in handles:
                h.remove()
            bank.begin_forward()  # drop the kept-alive inputs
            world, rank = _dist_world(), _dist_rank()
            # distinct Hessians in order of first use, each owned by one rank (all of them by rank 0 when world == 1)
            slot_order: List[int] = []
            for lname in layers:
                sl = bank.layer_to_slot[lname]
                if sl not in slot_order:
                    slot_order.append(sl)
            owners_by_pos = slot_owners(len(slot_order), world)
            owner_of_slot = {sl: owners_by_pos[i] for i, sl in enumerate(slot_order)}
            for sl in range(len(bank.acc)):
                owner_of_slot.setdefault(sl, 0)
            nsamples = self._global_nsamples(bank.nsamples)
            if world > 1:
                bank.reduce_to_owners(owner_of_slot)
            t0 = self._sync_time("hessian_fwd", t0)
            # ---- Hessian -> inverse factor, once per distinct input, on its owner rank (gptq.py:1189-1231) ----
            # The distinct Hessians of a block (4 for Llama) and then its layers (7) are independent given the
            # Hessians: the K2 chains (latency-bound diagonal-tile factorisations) and the column loops (N/4 warps each,
            # far from filling 148 SMs) are issued round-robin on side streams and joined before forward #2.  With
            # several ranks the factor is broadcast right before the first layer that needs it, so the big (down_proj)
            # factorisation on its owner overlaps the other layers' column loops on every rank.
            results = {}
            by_slot: Dict[tuple, dict] = {}
            keys_of_slot: Dict[int, set] = {}
            main = torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
            n_side = 1 if (self.profile or main is None) else max(1, int(os.environ.get("B200WOQ_GPTQ_STREAMS", "4")))
            side = self._side_streams(n_side) if n_side > 1 else []
            for st in side:
                st.wait_stream(main)

            def on(idx):
                return torch.cuda.stream(side[idx % len(side)]) if side else contextlib.nullcontext()

            for lname in layers:
                cfg = self.get_layer_config(self.full_name(lname, block_idx))
                keys_of_slot.setdefault(bank.layer_to_slot[lname], set()).add(
                    (float(cfg["percdamp"]), _order_tag(cfg)))
            for lname, layer in layers.items():
                cfg = self.get_layer_config(self.full_name(lname, block_idx))
                if cfg.get("static_groups") and cfg["group_size"] not in (-1, layer.weight.shape[-1]):
                    # the reference itself cannot run this: with static_groups its fasterquant returns ONE scale column
                    # (gptq.py:1193-1200 skips the per-group append, :1339-1341 appends the last quantizer only) and the
                    # export then raises IndexError (utility.py:483-537 indexes scale[:, i]); verified on the live reference
                    raise NotImplementedError("static_groups with group_size < in_features: the reference raises "
                                              "IndexError at export (gptq.py:1339-1341); nothing to be drop-in for")
                if cfg.get("fp8_aware"):
                    raise NotImplementedError("fp8_aware is the reference's W4A8 path for HPU (Gaudi) fp8 matmuls: out of scope")
                if cfg.get("hybrid_order") and cfg.get("act_order"):
                    raise AssertionError("Error: hybrid_act_order is not allowed with act_order")   # gptq.py:1204
                if cfg.get("use_double_quant") and str(cfg.get("double_quant_dtype", "int")) != "int":
                    raise NotImplementedError("double quant of scales: int dtype only")
                slot = bank.layer_to_slot[lname]
                key = (slot, float(cfg["percdamp"]), _order_tag(cfg))
                if key in by_slot:
                    continue
                ent = dict(Hinv=None, dead=None, perm=None, info=None, done=None, shared=world == 1,
                           owner=owner_of_slot[slot], C=bank.acc[slot].shape[0], act_order=bool(_order_tag(cfg)))
                if ent["owner"] == rank:
                    with on(len(by_slot)):
                        # finalize is in place: clone only when several configs share one raw accumulator
                        Hc = bank.acc[slot].clone() if len(keys_of_slot[slot]) > 1 else bank.acc[slot]
                        Hc, dead = ops.hessian_finalize(Hc, nsamples, cfg["percdamp"])
                        perm = None
                        if cfg["act_order"]:
                            perm = torch.argsort(torch.diag(Hc), descending=True)
                        elif cfg.get("hybrid_order"):
                            perm = hybrid_order_perm(torch.diag(Hc), int(cfg["group_size"]))
                        if perm is not None:
                            Hc = Hc[perm][:, perm].contiguous()
                            dead = dead[perm].contiguous()
                        t1 = time.perf_counter()
                        info = torch.zeros(1, dtype=torch.int32, device=self.device)
                        Hinv = ops.cholesky_inverse_upper(Hc, info=info, check=False)
                        self._sync_time("cholesky", t1)
                        done = torch.cuda.Event() if side else None
                        if done is not None:
                            done.record()
                    ent.update(Hinv=Hinv, dead=dead, perm=perm, info=info, done=done)
                by_slot[key] = ent
            fq_side = side
            for li, (lname, layer) in enumerate(layers.items()):
                cfg = self.get_layer_config(self.full_name(lname, block_idx))
                key = (bank.layer_to_slot[lname], float(cfg["percdamp"]), _order_tag(cfg))
                ent = by_slot[key]
                # ---- fasterquant (gptq.py:704-713) ----
                t1 = time.perf_counter()
                with (torch.cuda.stream(fq_side[li % len(fq_side)]) if fq_side else contextlib.nullcontext()):
                    if not ent["shared"]:
                        self._share_factor(ent, ent["owner"], ent["C"], ent["act_order"])
                    elif fq_side and ent["done"] is not None:
                        torch.cuda.current_stream(self.device).wait_event(ent["done"])
                    Hinv, dead, perm = ent["Hinv"], ent["dead"], ent["perm"]
                    W = layer.weight.data
                    W = (W.t() if _is_conv1d(layer) else W).float()
                    W = (W[:, perm] if perm is not None else W).contiguous().clone()
                    r = self._fasterquant_rows_sharded(W, Hinv, dead, cfg)
                    if perm is not None:
                        inv = torch.argsort(perm)
                        r["Q"] = r["Q"][:, inv].contiguous()
                        r["codes"] = r["codes"][:, inv].contiguous()
                        if cfg["act_order"]:
                            r["perm"] = perm
                        else:
                            # hybrid_order (gptq.py:1320-1328): the groups were visited in `order`; put their
                            # parameters back in storage order.  order[j] = group processed j-th = perm[j*g] // g
                            g = int(cfg["group_size"])
                            inv_order = torch.argsort(perm[::g] // g)
                            r["scale"] = r["scale"][:, inv_order].contiguous()
                            r["zero"] = r["zero"][:, inv_order].contiguous()
                    Q = r.pop("Q")
                    layer.weight.data = (Q.t().contiguous() if _is_conv1d(layer) else Q).to(layer.weight.dtype)
                results[lname] = r
                logger.info(f"block {block_idx} {lname}: error {r['losses'].sum().item():.6f}" if self.profile else
                            f"block {block_idx} {lname} quantized")
                self._sync_time("fasterquant", t1)
            infos = [e["info"] for e in by_slot.values() if e["info"] is not None]
            for st in fq_side:
                main.wait_stream(st)
            del bank, by_slot
            # ---- forward pass #2: propagate quantised outputs (gptq.py:749-762) ----
            t0 = time.perf_counter()
            batch_num = self.cache_kwargs.pop("batch_num")
            for j in range(batch_num):
                args, kw = self._batch(j)
                out = self._hidden(block(*args, **kw))
                if "hidden_states" in self.cache_kwargs:
                    self.cache_kwargs["hidden_states"][j] = out
                else:
                    self.cache_args[0][j] = out
            self.cache_kwargs["batch_num"] = batch_num
            # factorisation status of this group's Hessians: the reference's torch.linalg.cholesky raises on a
            # non-positive-definite H (gptq.py:1228).  The status is copied to pinned host memory asynchronously and
            # examined when the NEXT block starts (or at the end of the run), so no block ever waits on the host.
            if infos:
                self._queue_factor_status(infos, block_idx)
            t0 = self._sync_time("propagate", t0)
            # ---- export: pack on the device (gptq.py:769-849) ----
            for lname, layer in layers.items():
                cfg = self.get_layer_config(self.full_name(lname, block_idx))
                r = results[lname]
                if _is_conv1d(layer):
                    in_f, out_f = layer.weight.shape[0], layer.weight.shape[1]
                else:
                    in_f, out_f = layer.in_features, layer.out_features
                new_module = B200WeightOnlyLinear(in_f, out_f, dtype=cfg["dtype"], bits=cfg["bits"],
                                                  group_size=cfg["group_size"], zp=not cfg["sym"],
                                                  bias=layer.bias is not None, g_idx=r.get("perm") is not None,
                                                  device=self.device)
                new_module.pack_stored(r["codes"], r["scale"], None if cfg["sym"] else r["zero"], layer.bias,
                                       g_idx=r.get("perm"))
                set_module(block, lname, new_module)
            self._sync_time("pack", t0)
        if fetched:
            # not ours: the owner keeps the packed block, this rank only needed it for its share of the calibration
            from ..utils import sharded

            # (stream-safe without a host sync: the weights were allocated on the main stream and every side stream that
            # read them has been joined into it before forward #2 was queued)
            blocks[block_idx] = sharded.release_block(block)
            return blocks[block_idx]
        if self.offload_packed_to_host:
            block = block.to("cpu")
            blocks[block_idx] = block
        return block


def _order_tag(cfg) -> str:
    """Which column order the layer's column loop runs in: "" (natural), "act" (act_order, gptq.py:1212-1216) or
    "hybrid:<g>" (hybrid_order, :1203-1209).  Part of the key under which a factorised Hessian is shared."""
    if cfg.get("act_order"):
        return "act"
    if cfg.get("hybrid_order"):
        return f"hybrid:{int(cfg['group_size'])}"
    return ""


def hybrid_order_perm(diag_h: torch.Tensor, group_size: int) -> torch.Tensor:
    """gptq.py:1389-1461 (`compute_local_perms` / `compute_global_perm` / `compose_final_perm`): columns are sorted by
    descending diag(H) INSIDE their group and the groups by their largest diagonal entry -- the salient columns come
    first like with act_order, but no column leaves its group, so the packed module needs no g_idx."""
    C = diag_h.numel()
    if group_size <= 0 or C % group_size:
        raise NotImplementedError("hybrid_order needs group_size > 0 dividing in_features (the reference drops the tail)")
    d = diag_h.view(C // group_size, group_size)
    local = torch.argsort(d, dim=1, descending=True)
    order = torch.argsort(d.max(dim=1).values, descending=True)
    return (local[order] + (order * group_size).view(-1, 1)).reshape(-1)


class GPTQuantizer(Quantizer):
    """gptq.py:1651-1716."""

    def __init__(self, quant_config=None):
        super().__init__(quant_config if quant_config is not None else {})

    @torch.no_grad()
    def prepare(self, model, nsamples=128, max_seq_length=2048, use_max_length=True, device=None, use_layer_wise=False,
                model_path=None, quant_lm_head=False, use_block_wise=False, *args, **kwargs):
        assert isinstance(model, torch.nn.Module), "only support torch module"
        self.model_device = get_model_device(model)
        self.gptq_quantizer = RAWGPTQuantizer(model, weight_config=self.quant_config, nsamples=nsamples,
                                              use_max_length=use_max_length, max_seq_length=max_seq_length,
                                              use_layer_wise=use_layer_wise, model_path=model_path,
                                              quant_lm_head=quant_lm_head, use_block_wise=use_block_wise)
        self.gptq_quantizer.prepare_for_calibration()
        return self.gptq_quantizer.model

    @torch.no_grad()
    def convert(self, model, *args, **kwargs):
        self.gptq_quantizer.model = model
        self.gptq_quantizer.remove_prepare_for_calibration()
        q_model = self.gptq_quantizer.execute_quantization()
        # packed modules live on the B200; keep the rest of the model with them.  A layer-sharded model (utils/sharded.py)
        # was built on the device already and its remote blocks are storage-less skeletons that cannot be moved
        shard = getattr(q_model, "_b200_shard", None)
        if shard is None or shard["world"] == 1:
            q_model = q_model.to(self.gptq_quantizer.device)
        logger.info("GPTQ quantizing done.")
        return q_model
