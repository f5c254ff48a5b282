"""Drop-in CLI surface (-m gpu): `python -m mmlf.train.cli` / `python -m mmlf.validate.cli` through the `mmlf` shim package
with the five flags the north star names (--train_shift, --model_uncert, --model_discrete, --val_ensamble,
--train_loss_multimodal), on the synthetic light-field source, a few iterations each."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mod, *args):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, '-m', mod, *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + '\n' + r.stderr[-4000:]
    return r.stdout


COMMON = ['--model_chs', '16', '--train_bs', '4', '--train_ps', '32', '--train_lr', '1e-3', '--train_num_workers', '0',
          '--val_interval', '2', '--max_iterations', '3', '--train_warm_start']


@pytest.mark.parametrize('flags,val_flags', [
    (['--train_shift', '2.5', '--model_uncert', '--val_ensamble', '--val_disp_step', '1.0'],
     ['--val_ensamble', '--val_disp_step', '1.0', '--train_shift', '2.5']),
    (['--model_discrete', '--train_loss_multimodal'], ['--model_discrete']),
    (['--model_cross', '--train_eval_mode', '--train_eval_mode_start', '1'], []),      # iterations 1, 2 train in eval() mode
    (['--gpu_augment', '--train_shift', '2.5', '--train_max_downscale', '2'], []),
])
def test_train_then_validate_cli(tmp_path, flags, val_flags):
    out = str(tmp_path)
    log = _run('mmlf.train.cli', out, *COMMON, *flags)
    lines = [l for l in log.splitlines() if l.strip() and l.strip()[0].isdigit()]
    assert len(lines) == 3, log
    for l in lines:
        loss = float(l.split(',')[1])
        assert loss == loss and abs(loss) < 1e4, l              # finite
    state = torch.load(os.path.join(out, 'checkpoint.pt'), map_location='cpu')
    assert set(state) == {'model_state_dict', 'optimizer_state_dict', 'hyper_parameters', 'epoch', 'iteration', 'loss'}
    assert os.path.exists(os.path.join(out, 'log.csv'))
    val = _run('mmlf.validate.cli', out, out, '--size', '48', *val_flags)
    assert 'MSE & BadPix007' in val
    # resume from the checkpoint just written
    _run('mmlf.train.cli', out, *COMMON, *flags, '--train_resume')
