"""transformer_b200 -- B200-native drop-in for kuetuofa/Transformer's training hot path.

Public names mirror the reference's modules: Transformer (Transformer.py), Train / CustomSchedule (train.py),
DistributedTrain (distributed_train.py), transformer_flags / flags_dict (utils.py).
"""
__version__ = "0.1.0"
