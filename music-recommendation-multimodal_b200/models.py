"""Reference-shaped modules over the B200 engine.

Same class names, constructor keyword arguments, ``forward`` signatures and state-dict keys as
the reference (src/models/two_tower.py:8-168, user_tower.py:4-144, item_tower.py:100-152), with
the four modality encoders out of scope: ``batch['target_audio'|'target_image'|
'target_input_ids'|'target_tabular']`` (and the ``audio``/``images``/``input_ids``/``tabular``
arguments) carry precomputed (B, 128) embeddings.

Every parameter is a view into the engine's flat fp32 buffer, so ``state_dict()`` /
``load_state_dict()`` use the reference layout and the fused AdamW updates the same storage.
``TwoTowerModel.forward`` is differentiable through a single autograd node whose backward is the
hand-written CUDA backward pass (for stock optimizers); the fast path (`train.train_one_epoch`
with `FusedAdamW`) skips autograd entirely.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .engine import TwoTowerEngine
from .synthetic import TwoTowerConfig


class _ParamTree(nn.Module):
    """Nested container that exposes engine views under dotted reference names."""

    def _insert(self, dotted: str, tensor: torch.Tensor, buffer: bool = False) -> None:
        node = self
        *path, leaf = dotted.split(".")
        for part in path:
            if part not in node._modules:
                node.add_module(part, _ParamTree())
            node = node._modules[part]
        if buffer:
            node.register_buffer(leaf, tensor)
        else:
            node.register_parameter(leaf, nn.Parameter(tensor, requires_grad=True))


class SequentialUserEncoder(_ParamTree):
    """SASRec user tower (src/models/user_tower.py:4-144)."""

    def __init__(self, vocab_size: int = 0, num_genders: int = 1, num_countries: int = 1, embedding_dim: int = 256,
                 max_seq_len: int = 50, num_heads: int = 4, num_layers: int = 2, dropout: float = 0.1,
                 _engine: Optional[TwoTowerEngine] = None):
        super().__init__()
        if _engine is None:
            cfg = TwoTowerConfig(vocab_size=vocab_size, num_genders=num_genders, num_countries=num_countries,
                                 max_seq_len=max_seq_len, embedding_dim=embedding_dim, num_heads=num_heads,
                                 num_layers=num_layers, dropout=dropout)
            _engine = TwoTowerEngine(cfg)
            from .synthetic import make_state_dict
            _engine.load_state_dict(make_state_dict(cfg, seed=torch.initial_seed() % (2 ** 31), perturb=0.0))
        object.__setattr__(self, "_eng", _engine)
        self.embedding_dim, self.max_seq_len = _engine.cfg.embedding_dim, _engine.cfg.max_seq_len
        for name, t in _engine.p.items():
            if name.startswith("user_tower."):
                self._insert(name[len("user_tower."):], t)

    def forward(self, history_ids, user_gender, user_country, history_mask=None):
        eng = self._eng
        B, L = history_ids.shape
        ws = eng.workspace(B, L)
        if not eng.shadow_valid:
            eng.refresh_shadow()
        mask = None if history_mask is None else history_mask.long().contiguous()
        eng.user_forward(ws, history_ids.contiguous(), mask, user_gender.contiguous(), user_country.contiguous(),
                         training=self.training)
        return ws["u"].clone()


class MultimodalItemEncoder(_ParamTree):
    """Late-fusion item tower (src/models/item_tower.py:100-152), fusion part only."""

    def __init__(self, tabular_input_dim: int = 128, embedding_dim: int = 256, audio_dim: int = 128,
                 visual_dim: int = 128, text_model_name: str = "microsoft/mdeberta-v3-base", text_dim: int = 128,
                 tabular_dim: int = 128, use_lora: bool = True, _engine: Optional[TwoTowerEngine] = None):
        super().__init__()
        assert _engine is not None, "construct through TwoTowerModel"
        object.__setattr__(self, "_eng", _engine)
        for name, t in _engine.p.items():
            if name.startswith("item_tower."):
                self._insert(name[len("item_tower."):], t)
        self._insert("fusion_layer.1.running_mean", _engine.bn_running_mean, buffer=True)
        self._insert("fusion_layer.1.running_var", _engine.bn_running_var, buffer=True)
        self._insert("fusion_layer.1.num_batches_tracked", _engine.bn_num_batches, buffer=True)

    def forward(self, images, audio, input_ids, attention_mask, tabular):
        eng = self._eng
        B = audio.shape[0]
        ws = eng.item_workspace(B)
        if not eng.shadow_valid:
            eng.refresh_shadow()
        f = [t.float().contiguous() for t in (audio, images, input_ids, tabular)]
        eng.item_forward(ws, f[0], f[1], f[2], f[3], training=self.training)
        # un-normalised fusion output = LayerNorm(y2): recover it from the normalised one is lossy,
        # so expose the normalised embedding's pre-image through the same chain without L2
        from . import ops
        out = torch.empty_like(ws["in"])
        p = eng.p
        ops.chain_fwd(ws["y2"], ln=(p["item_tower.fusion_layer.5.weight"], p["item_tower.fusion_layer.5.bias"]),
                      out_f32=out)
        return out


class _TwoTowerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, batch, *params):
        eng = model.engine
        loss, logits, u, i = eng.forward(batch, training=model.training)
        ctx.model = model
        ctx.names = list(eng.p.keys())
        return loss.clone(), logits.clone(), u.clone(), i.clone()

    @staticmethod
    def backward(ctx, dloss, dlogits, du, di):
        eng = ctx.model.engine
        eng.grad.zero_()
        eng.backward(loss_scale=float(dloss.item()))
        grads = tuple(eng.g[n].clone() for n in ctx.names)
        # gradients handed to autograd mean an external optimizer is about to change the fp32 masters behind
        # the bf16 operand shadow: the next forward must re-cast it
        eng.shadow_valid = False
        return (None, None) + grads


class TwoTowerModel(nn.Module):
    """Drop-in for the reference TwoTowerModel (src/models/two_tower.py:8-168)."""

    def __init__(self, vocab_size: int, tabular_input_dim: int = 128, num_genders: int = 1, num_countries: int = 1,
                 max_seq_len: int = 50, user_embedding_dim: int = 256, user_num_heads: int = 4,
                 user_num_layers: int = 2, user_dropout: float = 0.1, item_embedding_dim: int = 256,
                 audio_dim: int = 128, visual_dim: int = 128, text_model_name: str = "microsoft/mdeberta-v3-base",
                 text_dim: int = 128, tabular_dim: int = 128, use_lora: bool = True, temperature: float = 0.07,
                 device=None, seed: int = 0):
        super().__init__()
        assert user_embedding_dim == item_embedding_dim, \
            f"User dim ({user_embedding_dim}) must match Item dim ({item_embedding_dim})"
        assert audio_dim == visual_dim == text_dim == tabular_dim == 128
        cfg = TwoTowerConfig(vocab_size=vocab_size, num_genders=num_genders, num_countries=num_countries,
                             max_seq_len=max_seq_len, embedding_dim=user_embedding_dim, num_heads=user_num_heads,
                             num_layers=user_num_layers, dropout=user_dropout, temperature=temperature)
        self.temperature = temperature
        dev = None if device is None else torch.device(device)
        engine = TwoTowerEngine(cfg, dev)
        from .synthetic import make_state_dict
        engine.load_state_dict(make_state_dict(cfg, seed=seed, perturb=0.0))   # reference init scales
        object.__setattr__(self, "engine", engine)
        self.user_tower = SequentialUserEncoder(_engine=engine)
        self.item_tower = MultimodalItemEncoder(_engine=engine)
        engine.rebind_hooks.append(self._rebind_parameters)

    def _rebind_parameters(self) -> None:
        """The engine moved its flat buffers (into a symmetric arena for multi-GPU training): Parameters follow."""
        for name, prm in self.named_parameters():
            prm.data = self.engine.p[name]

    # -- checkpoints in the reference layout (src/train.py:327-330; loaders strip 'module.')
    def load_state_dict(self, state_dict, strict: bool = False):  # type: ignore[override]
        """Reference checkpoints also carry the four modality encoders (out of scope here): their keys are
        reported as unexpected, never an error. A key of the hot path that is absent is an error under
        ``strict`` and is reported (and left at its current value) otherwise."""
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        own = set(self.engine.p) | {"item_tower.fusion_layer.1." + b for b in
                                    ("running_mean", "running_var", "num_batches_tracked")}
        missing = sorted(k for k in self.engine.p if k not in sd)
        unexpected = sorted(k for k in sd if k not in own)
        if missing and strict:
            raise RuntimeError(f"Error(s) in loading state_dict for TwoTowerModel: missing keys {missing}")
        self.engine.load_state_dict({**{k: self.engine.p[k] for k in missing}, **sd})
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _param_version(self) -> int:
        return sum(p._version for p in self.parameters())

    def _sync_shadow(self) -> None:
        """The bf16 operand shadow follows the fp32 masters: any in-place change of a parameter since the last
        forward (a stock optimizer's step, ``p.data`` edits, ``p.copy_``) bumps its version counter and forces a
        re-cast; the fused AdamW refreshes the shadow itself."""
        v = self._param_version()
        if v != getattr(self, "_seen_version", None):
            self.engine.shadow_valid = False
            object.__setattr__(self, "_seen_version", v)

    def _ordered_params(self):
        named = dict(self.named_parameters())
        return tuple(named[n] for n in self.engine.p)

    def _batch(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = {}
        for k, v in batch.items():
            if not isinstance(v, torch.Tensor):
                continue
            v = v.to(self.engine.device)
            out[k] = (v.float() if v.is_floating_point() else v.long()).contiguous()
        return out

    def forward(self, batch: Dict[str, torch.Tensor]):
        b = self._batch(batch)
        self._sync_shadow()
        params = self._ordered_params()
        if torch.is_grad_enabled() and self.training:
            return _TwoTowerFunction.apply(self, b, *params)
        loss, logits, u, i = self.engine.forward(b, training=self.training)
        return loss.clone(), logits.clone(), u.clone(), i.clone()

    def get_user_embedding(self, history_ids, history_mask=None, user_gender=None, user_country=None):
        """Normalised user embedding (src/models/two_tower.py:144-157)."""
        eng = self.engine
        ids = history_ids.to(eng.device).long().contiguous()
        if user_gender is None:
            user_gender = torch.zeros_like(ids[:, 0])
        if user_country is None:
            user_country = torch.zeros_like(ids[:, 0])
        mask = None if history_mask is None else history_mask.to(eng.device).long().contiguous()
        B, L = ids.shape
        ws = eng.workspace(B, L)
        self._sync_shadow()
        if not eng.shadow_valid:
            eng.refresh_shadow()
        u = eng.user_forward(ws, ids, mask, user_gender.to(eng.device).long().contiguous(),
                             user_country.to(eng.device).long().contiguous(), training=self.training)
        return u.clone()

    def get_item_embedding(self, images, audio, input_ids, attention_mask, tabular):
        """Normalised item embedding (src/models/two_tower.py:159-168)."""
        eng = self.engine
        f = [t.to(eng.device).float().contiguous() for t in (audio, images, input_ids, tabular)]
        ws = eng.item_workspace(f[0].shape[0])
        self._sync_shadow()
        if not eng.shadow_valid:
            eng.refresh_shadow()
        return eng.item_forward(ws, f[0], f[1], f[2], f[3], training=self.training).clone()
