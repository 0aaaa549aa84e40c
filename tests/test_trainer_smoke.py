"""the loss under its real caller.  tools/trainer_smoke.py restates AMTrainer.forward
(cat/ctc/train.py:172-197), the unified trainer's second criterion call per step (cat/ctc/train_unified.py:248,267) and the
manager's step (cat/shared/manager.py:524-547) -- the reference package itself cannot be imported offline (`import jieba`).
1 GPU here; the DDP x N run (mp.spawn + NCCL, cat/shared/coreutils.py:493-504) is `python tools/trainer_smoke.py --gpus N`,
logged under profiles/."""
import os
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("flags", [[], ["--amp"], ["--from-logits", "--amp"], ["--no-unified"]])
def test_trainer_smoke_single_gpu(flags):
    import trainer_smoke
    log = trainer_smoke.run(trainer_smoke.parse(["--steps", "30"] + flags))
    assert len(log) == 30
    assert sum(log[-5:]) < sum(log[:5])            # (run() itself raises unless the loss fell by >= 10 %)
