"""Glue a reference maintainer uses to put the accelerated query pass behind the reference's own driver
(scripts/Run_active_learning.py:165-173) while training and evaluation stay reference code.

    ref = RefAL(cfg, opt)                                   # the reference class builds datasets / estimator / AE
    al = vatlq.ActiveLearning(cfg, opt, **integration.from_reference(ref))
    while True:
        al.eval_and_query()
        if al.outcome() is not None: break

`from_reference(ref)` injects the reference's estimator, loader and auto-encoder and three hooks:
  * eval_hook     collects, per batch, the prediction records the reference builds in its per-person loop
                  (ActiveLearning.py:309-326: bbox, image_id, id, score, keypoints, GT keypoints)
  * metrics_hook  writes them as the reference does and calls the reference's evaluate_mAP / ospa_for_loc
                  (:438-447) — evaluation stays reference code
  * retrain_hook  copies the query-side state into the reference object (ids, mOKS, round) and calls the
                  reference's retrain_model() with the epoch rule of outcome() (:179-185, 651-686)
Nothing here computes scores or selections; it only moves state between the two objects.
"""
from __future__ import annotations

import copy
import json
import os

import numpy as np


def sync_to_reference(ref, al):
    """Copy what the reference's retrain_model / outcome read from `self` (ActiveLearning.py:166-205,651-686)."""
    IC = type(ref.labeled_id)                        # the reference's alipy IndexCollection
    ref.round_cnt = al.round_cnt
    ref.labeled_id = IC(list(al.labeled_id.index))
    ref.unlabeled_id = IC(list(al.unlabeled_id.index))
    ref.retrain_id = IC(list(al.retrain_id.index))
    ref.moks_queried = al.moks_queried
    ref.query_size = al.query_size
    ref.is_early_stop = al.is_early_stop
    return ref


def retrain_with(ref):
    """retrain_hook: the reference retrains its estimator (and re-fits the auto-encoder) on the ids we selected."""
    def hook(al):
        sync_to_reference(ref, al)
        cfg = ref.cfg
        if not getattr(ref, "continual", True):                                                   # (:179-182)
            ref.model, ref.optimizer, ref.scheduler = ref.initialize_estimator()
            ref.retrain_epoch = int(cfg.RETRAIN.BASE * len(ref.labeled_id.index) / len(ref.eval_dataset)
                                    + cfg.RETRAIN.ALPHA * (1 - ref.moks_queried))
        else:                                                                                     # (:183)
            ref.retrain_epoch = int(cfg.RETRAIN.ALPHA * (1 - ref.moks_queried))
        ref.retrain_model()                                                                       # (:185, 651-686)
        al.model = ref.model
        if hasattr(ref, "AE"):
            al.AE = ref.AE           # re-initialised and fine-tuned after every retrain (:681-685): re-packed at the next query
    return hook


class PredictionCollector:
    """eval_hook + metrics_hook: the json records of ActiveLearning.py:309-326 and the evaluation of :438-447."""

    def __init__(self, ref, evaluate_mAP=None, ospa_for_loc=None, bbox_xyxy_to_xywh=None):
        self.ref = ref
        if evaluate_mAP is None:
            from alphapose.utils.metrics import evaluate_mAP          # reference code
        if ospa_for_loc is None:
            from JRDB_toolkit.pose_eval import ospa_for_loc           # reference code (ActiveLearning.py imports it)
        if bbox_xyxy_to_xywh is None:
            from alphapose.utils.bbox import bbox_xyxy_to_xywh        # reference code
        self.evaluate_mAP, self.ospa_for_loc, self.to_xywh = evaluate_mAP, ospa_for_loc, bbox_xyxy_to_xywh
        self.kpt_json, self.kpt_json_ann, self.GT_json = [], [], []

    def eval_hook(self, al, batch, kpts):
        idxs, GTkpts, img_ids, ann_ids, bboxes_ann = batch[0], batch[4], batch[5], batch[6], batch[8]
        k = kpts.detach().float().cpu().numpy().reshape(len(idxs), -1)
        for j in range(len(idxs)):
            scores = k[j, 2::3]
            gt = np.asarray(GTkpts[j]).reshape(-1).tolist()
            data = {"bbox": self.to_xywh(np.asarray(bboxes_ann[j]).tolist()), "image_id": int(img_ids[j]), "id": int(ann_ids[j]),
                    "score": float(np.mean(scores) + 1.25 * np.max(scores)), "category_id": 1,
                    "keypoints": k[j].tolist(), "GT_keypoints": gt}
            self.kpt_json.append(data)
            data_ann = copy.deepcopy(data)
            if int(idxs[j]) in al.labeled_id:
                data_ann["keypoints"] = gt
            self.kpt_json_ann.append(data_ann)
            data_GT = copy.deepcopy(data)
            data_GT["keypoints"] = gt
            self.GT_json.append(data_GT)

    def metrics_hook(self, al, kpts_all):
        ref, wd = self.ref, self.ref.opt.work_dir
        if al.OKS_dict is not None:
            for rec, coll in ((self.kpt_json, None), (self.kpt_json_ann, None), (self.GT_json, None)):
                for pos, d in enumerate(rec):
                    d["OKS"] = al.OKS_dict.get(pos)
        gt_path = ref.save_GT_dict(self.GT_json)                                                   # (:439, 692-705)
        out = {}
        for name, rec, key_ap, key_ospa in (("predicted_kpt.json", self.kpt_json, "res", "ospa"),
                                            ("predicted_kpt_ann.json", self.kpt_json_ann, "res_ann", "ospa_ann")):
            path = os.path.join(wd, name)
            with open(path, "w") as fid:
                json.dump(rec, fid)
            out[key_ap] = self.evaluate_mAP(path, ann_type="keypoints", ann_file=gt_path)          # (:442, 445)
            out[key_ospa] = self.ospa_for_loc(ann_json_path=gt_path, pr_json_path=path)            # (:443, 447)
        self.kpt_json, self.kpt_json_ann, self.GT_json = [], [], []
        return out


def from_reference(ref, with_metrics: bool = True, **collector_kwargs) -> dict:
    """Keyword arguments for vatlq.ActiveLearning(cfg, opt, **from_reference(ref))."""
    kw = dict(model=ref.model, eval_loader=ref.eval_loader, eval_len=ref.eval_len, AE=getattr(ref, "AE", None),
              retrain_hook=retrain_with(ref))
    if with_metrics:
        pc = PredictionCollector(ref, **collector_kwargs)
        kw.update(eval_hook=pc.eval_hook, metrics_hook=pc.metrics_hook)
    return kw
