#!/usr/bin/env python
"""Multi-attribute walk training - drop-in for the reference's ``train_multi_attr.py`` (BASELINE config 4) over the
B200-native hot path.

    python train_multi_attr.py --model stylegan_v2_real --transform scene --num_samples 20000 --learning_rate 1e-4 \\
        --latent w --attrList night,dark --walk_type linear --loss l2 --overwrite_config --prefix DarkNight \\
        --models_dir ./models_scene_multi_attr --no_gan_loss --no_content_loss [--walk_mlp 1 --size 1024 --batch_size 16]
    torchrun --nproc-per-node 8 train_multi_attr.py ...

Same loop as ``train.py`` (``train.train_loop``) with the semantics the reference script intends but cannot run as
shipped (it unpacks two values from the one-value ``get_alphas`` of ``stylegan_v2_real``): the
sampled alpha is a delta on the regressor's current prediction, targets are clamped to [0, 1]
(``train.multi_attr_targets``), three epochs by default, and ``loss_values.npy`` is written next to the checkpoints
(train_multi_attr.py:46-58, 109-116, 224-226).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import train


def main(argv=None):
    return train.main(argv, multi_attr=True)


if __name__ == "__main__":
    main()
