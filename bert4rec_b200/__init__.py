"""bert4rec_b200 -- B200-native (sm_100a) BERT4Rec masked-item training + ranking path.

Drop-in for the hot path of maneymarkus/BERT4Rec: the module layout mirrors ``bert4rec.*`` of the reference
(``models``, ``models.components.networks``, ``trainers``, ``trainers.optimizers``, ``evaluation``,
``dataloaders.samplers``, ``tokenizers``), tensors are torch tensors instead of tf tensors, and all arithmetic on the
path runs in hand-written CUDA kernels behind the C ABI of ``include/b4r.h`` (no CPU fallback).
"""
__version__ = "0.1.0"
