"""`custom_imports=dict(imports=['dsl_b200.plugin', 'dsl_b200.plugin_runner'])`: additionally replace
RUNNERS['SemiEpochBasedRunner'] (mmdet/runner/hooks/semi_epoch_based_runner.py:49) by the fused runner of
dsl_b200.runner, so that `tools/train.py` -> `train_detector` -> `build_runner(cfg.runner)` executes every iteration as
one captured CUDA-graph step. Kept apart from dsl_b200.plugin because it changes HOW an iteration runs, not only which
classes answer to the model registry keys."""
from . import plugin

REGISTERED = plugin.register(runner=True)
