"""``--network`` plug-ins: ``wseg_b200.network.resnet38_contrast`` is a drop-in for the reference's
``network.resnet38_contrast`` (contrast_train.py:40,60; contrast_infer.py:21,32)."""
