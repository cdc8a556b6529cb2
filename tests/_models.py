"""Builds this repository's AREAD module for a fixture spec (shared by CPU and GPU tests)."""
import importlib
import types

import numpy as np
import torch

from oracle import synth

PKG = importlib.import_module("aread-multi-domain-recommendation_b200")


def make_config(spec, use_atten=True):
    return types.SimpleNamespace(
        domain_size={"synth": [100] * spec.n_domain}, dataset_name="synth", use_dcn=True, use_atten=use_atten,
        n_cross_layers=spec.n_cross_layers, mmoe_n_expert=spec.n_expert, atten_embed_dim=64, att_head_num=2,
        att_layer_num=3, att_res=True)


def build_model(spec, device, dropout=0.0, deterministic=True, use_atten=True):
    mh = {"multi_hot_flag": list(spec.flag), "itemid_idx": spec.itemid_idx, "seq_maxlen": spec.seq_maxlen,
          "method": spec.method}
    model = PKG.AREAD(np.asarray(spec.one_hot_field_dims), spec.embed_dim, mh, n_tower=tuple(spec.n_tower),
                      n_domain=spec.n_domain, base_model="mmoe", expert_dims=tuple(spec.expert_dims),
                      tower_dims=tuple(tuple(t) for t in spec.tower_dims), domain_idx=spec.domain_idx,
                      device=torch.device(device), dropout=dropout, config=make_config(spec, use_atten))
    model.reset_for_mask_update()
    if deterministic:
        model.load_state_dict(synth.deterministic_state(spec, with_attention=use_atten), strict=True)
    return model.to(device)
