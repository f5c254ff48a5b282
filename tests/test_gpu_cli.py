"""Drop-in CLI surface (-m gpu): `python -m mmlf.train.cli` / `python -m mmlf.validate.cli` through the `mmlf` shim package
with the five flags the north star names (--train_shift, --model_uncert, --model_discrete, --val_ensamble,
--train_loss_multimodal), a few iterations each: on the synthetic light-field source (--synthetic) and on an on-disk
dataset in the HCI4D layout (PNG views + PFM ground truth) written by tests/_fixtures.py."""
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


COMMON = ['--synthetic', '--model_chs', '16', '--train_bs', '4', '--train_ps', '32', '--train_lr', '1e-3', '--train_num_workers', '0',
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
    val = _run('mmlf.validate.cli', out, out, '--synthetic', '--size', '48', *val_flags)
    assert 'MSE & BadPix007' in val
    # resume from the checkpoint just written
    _run('mmlf.train.cli', out, *COMMON, *flags, '--train_resume')


def _run_fail(mod, *args):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, '-m', mod, *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0, r.stdout[-2000:]
    return r.stdout + r.stderr


@pytest.mark.parametrize('flags,val_flags', [
    (['--model_uncert', '--train_shift', '1.5'], ['--train_shift', '1.5']),
    (['--train_no_data_augment'], []),
])
def test_train_then_validate_on_an_hci4d_directory(tmp_path, flags, val_flags):
    """README workflow of the reference on real files: train on DATA/additional, validate on DATA/training, results
    dumped by HCI4D.save_batch (validate/cli.py:313)."""
    import numpy as np
    import _fixtures as fx
    from mmlf_b200.utils import pfm
    data, out = tmp_path / 'data', tmp_path / 'out'
    out.mkdir()
    for j, name in enumerate(('antinous', 'boardgames', 'dishes')):
        fx.write_hci_scene(str(data / 'additional' / name), 20 + j, 112, 112)
    gts = {}
    for j, name in enumerate(('boxes', 'cotton')):
        _, gts[name] = fx.write_hci_scene(str(data / 'training' / name), 30 + j, 64, 64, with_mpi=(j == 0))
    args = [str(out), '--train_trainset', str(data / 'additional'), '--train_valset', str(data / 'training'),
            '--model_chs', '16', '--train_bs', '4', '--train_ps', '32', '--train_lr', '1e-3', '--train_max_downscale', '2',
            '--val_interval', '2', '--max_iterations', '3', *flags]
    log = _run('mmlf.train.cli', *args)
    assert 'Caching dataset "additional"' in log and 'Caching dataset "training"' in log
    lines = [l for l in log.splitlines() if l.strip() and l.strip()[0].isdigit()]
    assert len(lines) == 3, log
    state = torch.load(str(out / 'checkpoint.pt'), map_location='cpu')
    assert 'synthetic' not in state['hyper_parameters']
    val = _run('mmlf.validate.cli', str(out), str(data / 'training'), *val_flags)
    assert 'MSE & BadPix007' in val and 'Processing scene 1' in val
    shift = float(val_flags[1]) if val_flags else 0.0
    for name in ('boxes', 'cotton'):
        sd = out / 'scenes' / name
        for f in ('center.png', 'gt.png', 'diff.png', 'result.png', 'gt.pfm', 'result.pfm', 'view_h_0.png', 'view_d_8.png'):
            assert (sd / f).exists(), f
        assert (sd / 'uncert.pfm').exists() == ('--model_uncert' in flags)
        assert (out / 'ours' / 'disp_maps' / f'{name}.pfm').exists() and (out / 'ours' / 'runtimes' / f'{name}.txt').exists()
        res = pfm.load(str(sd / 'result.pfm'))
        assert res.shape == (64, 64) and np.isfinite(res).all()
        # gt.pfm is the ground truth AFTER the Shift transform (gt - train_shift), stored bottom-up like the input file
        assert np.array_equal(np.flip(pfm.load(str(sd / 'gt.pfm')), 0), gts[name] - np.float32(shift))
    # missing dataset directories are an error, not a silent switch to synthetic data
    msg = _run_fail('mmlf.train.cli', str(out), '--max_iterations', '1')
    assert 'is not a directory' in msg
    msg = _run_fail('mmlf.validate.cli', str(out), str(out / 'ours' / 'runtimes'))
    assert 'no scene directories' in msg
