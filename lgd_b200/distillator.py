"""The five META_ARCH wrappers of the reference (models/distillator.py:23-494): same class names, registered into
detectron2's META_ARCH_REGISTRY so the reference's `train.py` builds them unchanged (train.py:247-248).

The bodies are orchestration only -- student forward (detectron2 / cvpods code, out of scope) -> DynamicTeacher ->
student head on the teacher features -> distill_loss -- and keep the reference's call conventions towards the
student detectors. Only the two hot-path calls (`self.teacher(...)`, `self.distill_loss(...)`) differ from the
reference: they run on liblgd_b200."""
from .base_distillator import BaseDistillator
from .registry import META_ARCH_REGISTRY


class _DistillatorCommon(BaseDistillator):
    #: name of the kwarg carrying the student's targets into forward_teacher (distillator.py:60-61 vs :259-260)
    TARGET_KW = 'gt_targets'

    def __init__(self, cfg=None):
        super().__init__(cfg)
        self.flag_seg_map = cfg.MODEL.DISTILLATOR.LABEL_ENCODER.LOAD_LABELMAP

    def forward(self, batched_inputs, **kwargs):
        if self.training:
            losses, r_features, features, images, targets = self.forward_student(batched_inputs)
            losses_tea, _, features_tea, masks, inst_labels = self.forward_teacher(
                batched_inputs, images=images, r_features=r_features, features=features, **{self.TARGET_KW: targets})
            losses_distill = self.distill_loss({'stu': features, 'tea': features_tea}, images, batched_inputs, masks,
                                               inst_labels)
            losses.update(losses_tea)
            losses.update(losses_distill)
            return losses
        processed_results, r_features, features, images = self.forward_student(batched_inputs)
        return self._eval(processed_results, r_features, features, images, batched_inputs, **kwargs)

    def forward_student(self, batched_inputs, **kwargs):
        return self.student(batched_inputs)

    def forward_teacher(self, batched_inputs, **kwargs):
        images, r_features, features = kwargs['images'], kwargs['r_features'], kwargs['features']
        features_tea, inst_labels, masks = self.teacher((batched_inputs, images, r_features, features))
        losses_tea = self._head_losses(features_tea, kwargs[self.TARGET_KW], images, batched_inputs)
        losses_tea = {k + '.tea': v for k, v in losses_tea.items()}
        return losses_tea, None, features_tea, masks, inst_labels

    #: run the student's FCOS-family head on the TEACHER features with lgd_b200.heads.FCOSHeadB200 (same parameters,
    #: tcgen05 convolutions + the library's GroupNorm(32) kernels on the NHWC teacher pyramid)
    #: when the head is the stock FCOSHead / POTOHead
    B200_HEAD = True

    def _fcos_predict_on_teacher(self, feats):
        """FCOSCT / ATSSCT / POTOCT.predict (customized_detectors/fcos.py:29-33) for the teacher features:
        (shifts, box_cls, box_delta[, box_center])"""
        head = getattr(self.student, 'head', None)
        if self.B200_HEAD and head is not None and feats[0].is_cuda and hasattr(self.student, 'shift_generator'):
            from .heads import FCOSHeadB200
            cache = self.__dict__.setdefault('_b200_head', {})      # not a submodule: the parameters stay the student's
            if cache.get('src') is not head:
                cache['src'] = head
                cache['head'] = FCOSHeadB200(head) if FCOSHeadB200.supports(head) else None
            if cache['head'] is not None:
                return (self.student.shift_generator(feats), *cache['head'](feats))
        return self.student.predict(feats)

    def _teacher_feature_list(self, batched_inputs, images, r_features, features, keys):
        features_tea, _, _ = self.teacher((batched_inputs, images, r_features, features))
        if isinstance(features_tea, dict):
            features_tea = [features_tea[f] for f in keys]
        return features_tea


@META_ARCH_REGISTRY.register()
class DistillatorRetinaNet(_DistillatorCommon):
    """models/distillator.py:23-114"""
    TARGET_KW = 'gt_labels_boxes'
    # B200_HEAD: lgd_b200.heads.RetinaNetHeadB200 when the head is the stock RetinaNetHead

    def _predict_on_teacher(self, feats):
        """RetinaNetCT.predict (customized_detectors/retinanet.py:36-45) for the teacher features"""
        head = getattr(self.student, 'head', None)
        if self.B200_HEAD and head is not None and feats[0].is_cuda and hasattr(self.student, 'anchor_generator'):
            from .heads import RetinaNetHeadB200
            cache = self.__dict__.setdefault('_b200_head', {})      # not a submodule: the parameters stay the student's
            if cache.get('src') is not head:
                cache['src'] = head
                cache['head'] = RetinaNetHeadB200.from_module(head) if RetinaNetHeadB200.supports(head) else None
            if cache['head'] is not None:
                logits, deltas = cache['head'](feats)
                return self.student.anchor_generator(feats), logits, deltas
        return self.student.predict(feats)

    def _head_losses(self, features_tea, targets, images, batched_inputs):
        gt_labels, gt_boxes = targets
        anchors, logits, deltas = self._predict_on_teacher([features_tea[f] for f in self.student.head_in_features])
        return self.student.losses(anchors, logits, gt_labels, deltas, gt_boxes)

    def _eval(self, processed_results, r_features, features, images, batched_inputs, **kwargs):
        feats = [features[f] for f in self.student.head_in_features]
        anchors, logits, deltas = self.student.predict(feats)
        if kwargs.get('eval_teacher', False):
            feats = self._teacher_feature_list(batched_inputs, images, r_features, features,
                                               self.student.head_in_features)
            anchors, logits, deltas = self.student.predict(feats)
        results = self.student.inference(anchors, logits, deltas, images.image_sizes)
        return self.student.get_processed_results(results, batched_inputs, images)


@META_ARCH_REGISTRY.register()
class DistillatorGeneralizedRCNN(_DistillatorCommon):
    """models/distillator.py:117-198 (Faster / Mask R-CNN students)"""
    TARGET_KW = 'gt_labels_boxes'

    def _head_losses(self, features_tea, gt_instances, images, batched_inputs):
        return self.student.predict(features_tea, images, gt_instances, batched_inputs)

    def _eval(self, processed_results, r_features, features, images, batched_inputs, **kwargs):
        if kwargs.get('eval_teacher', False):
            features_tea, _, _ = self.teacher((batched_inputs, images, r_features, features))
            return self.student.inference(batched_inputs, features=features_tea)[0]
        return processed_results


@META_ARCH_REGISTRY.register()
class DistillatorFCOS(_DistillatorCommon):
    """models/distillator.py:201-297"""

    def _head_losses(self, features_tea, targets, images, batched_inputs):
        gt_classes, gt_shifts_reg_deltas, gt_centerness = targets
        shifts, box_cls, box_delta, box_center = self._fcos_predict_on_teacher(
            [features_tea[f] for f in self.student.in_features])
        return self.student.losses(gt_classes, gt_shifts_reg_deltas, gt_centerness, box_cls, box_delta, box_center)

    def _eval(self, processed_results, r_features, features, images, batched_inputs, **kwargs):
        shifts, box_cls, box_delta, box_center = self.student.predict([features[f] for f in self.student.in_features])
        if kwargs.get('eval_teacher', False):
            feats = self._teacher_feature_list(batched_inputs, images, r_features, features, self.student.in_features)
            shifts, box_cls, box_delta, box_center = self.student.predict(feats)
        results = self.student.inference(box_cls, box_delta, box_center, shifts, images)
        return self.student.get_processed_results(results, batched_inputs, images)


@META_ARCH_REGISTRY.register()
class DistillatorPOTO(_DistillatorCommon):
    """models/distillator.py:299-395"""

    def _head_losses(self, features_tea, targets, images, batched_inputs):
        gt_classes, gt_shifts_reg_deltas = targets
        shifts, box_cls, box_delta = self._fcos_predict_on_teacher([features_tea[f] for f in self.student.in_features])
        return self.student.losses(gt_classes, gt_shifts_reg_deltas, box_cls, box_delta)

    def _eval(self, processed_results, r_features, features, images, batched_inputs, **kwargs):
        shifts, box_cls, box_delta = self.student.predict([features[f] for f in self.student.in_features])
        if kwargs.get('eval_teacher', False):
            feats = self._teacher_feature_list(batched_inputs, images, r_features, features, self.student.in_features)
            shifts, box_cls, box_delta = self.student.predict(feats)
        results = self.student.inference(box_cls, box_delta, shifts, images)
        return self.student.get_processed_results(results, batched_inputs, images)


@META_ARCH_REGISTRY.register()
class DistillatorATSS(DistillatorFCOS):
    """models/distillator.py:397-494 -- same head call convention as FCOS (centerness branch)."""
