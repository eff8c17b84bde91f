"""The `config["modules"]` tree BoxDreamer is constructed from (configs/model/transformer.yaml:10-71 with
configs/test.yaml's interpolations resolved), as a dict with attribute access (the reference reads both
`config["modules"]` and `self.dense_cfg.enable`, BoxDreamerModel.py:33,292)."""


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    def copy(self):
        return AttrDict({k: (v.copy() if isinstance(v, AttrDict) else v) for k, v in self.items()})


def make_config(img_size=224, num_layers=12) -> AttrDict:
    return AttrDict(modules=AttrDict(
        use_keypoints=False, use_matching=False, use_tracking=False, use_rgb=True, use_pp=True,
        ref_type="all", regression_intri=True, rotation_type=None, coordinate="object",
        pose_representation="bb8", bbox_representation="heatmap", patchify_rays=True,
        stage="decoder_only",
        dense_cfg=AttrDict(enable=False, filter_enable=True, filter="dino", filter_topk=5,
                           multi_round=False, sub_batch_size=5, fine_level=False, fine_topk=5,
                           dense_mem_friendly=False),
        decoder=AttrDict(d_model=768, nhead=8, num_decoder_layers=num_layers, camera_emb="MLP",
                         track_emb=None, match_emb=None, decoder_only=True, patch_size=14,
                         img_size=img_size, diff_emb=False, nvs_supervision=False,
                         ray_supervision=True, use_mask=False),
        tracker=AttrDict(ckpt_path=None, cfg=AttrDict(grid_size=20, freeze=True)),
        encoder=AttrDict(name="dino",
                         resnet=AttrDict(ckpt_path=None, cfg=AttrDict(model_type="resnet18", freeze=True)),
                         dino=AttrDict(ckpt_path=None, cfg=AttrDict(model_type="dinov2_vitb14_reg", freeze=True)),
                         spa=AttrDict(ckpt_path=None, cfg=AttrDict(model_type="spa_vit_base_patch16", freeze=True))),
    ))
