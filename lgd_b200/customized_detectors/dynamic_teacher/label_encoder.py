"""Parameter containers of the label encoder with the reference's state_dict names
(label_encoder.py:119-160, spatial_transformer.py:13-27). They only HOLD parameters: the forward/backward
arithmetic is engine.LabelEncoderTape (liblgd_b200 kernels). LayerNorms have no affine parameters."""
import torch.nn as nn


class STN(nn.Module):
    def __init__(self, k=64):
        super().__init__()
        self.conv1 = nn.Conv1d(k, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k * k)
        self.k = k


class LabelEncoder(nn.Module):
    def __init__(self, category_format='one_hot', box_format='x1y1x2y2', nr_fg_classes=80, add_context_box=False,
                 parse_mask=False):
        super().__init__()
        if category_format not in ('one_hot', 'norm_classes'):
            raise ValueError('category_format {} not supported yet !'.format(category_format))
        if box_format not in ('x1y1x2y2', 'x1y1wh'):
            raise ValueError('box_format {} not supported'.format(box_format))
        self.category_format, self.box_format = category_format, box_format
        self.nr_fg_classes, self.add_context_box = nr_fg_classes, add_context_box
        self.R, self.noise_std = 1, 0.0
        # label_encoder.py:136-141: one column class index / num_classes, or the one-hot class vector
        self.inp = 4 + (1 if category_format == 'norm_classes' else nr_fg_classes)
        self.parse_mask = parse_mask
        if parse_mask:   # LOAD_LABELMAP (Mask R-CNN recipe): + 7x7 mask descriptor (label_encoder.py:144-145)
            self.inp += 49
        self.stn_desc = STN(self.inp)
        self.stn_feat = STN(64)
        self.conv1 = nn.Conv1d(self.inp, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.conv4 = nn.Conv1d(1088, 256, 1)
