"""CPU tests of the host-side mirror of the reference interface (no kernels are launched)."""
import os

import numpy as np
import pytest
import torch

import _fixtures as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('variant,cross,n_keys,n_params', [
    ('base', False, 121, 4612166), ('upr', False, 121, 4613300), ('dpp', False, 121, 4778872),
    ('base', True, 94, 1208486)])
def test_state_dict_layout_matches_reference(variant, cross, n_keys, n_params):
    """Key set / shapes / parameter counts of SURVEY.md Appendix B (verified there on the reference)."""
    from mmlf_b200.model.feed_forward import FeedForward
    m = FeedForward(**fx.model_kwargs(variant, cross))
    sd = m.state_dict()
    assert len(sd) == n_keys
    assert sum(p.numel() for p in m.parameters()) == n_params
    assert sd['in_net_hv.0.0.weight'].shape == (70, 27, 2, 2)
    assert sd['in_net_hv.0.3.num_batches_tracked'].dtype == torch.int64
    assert ('in_net_id.0.0.weight' in sd) == (not cross)
    oc = {'base': 1, 'upr': 2, 'dpp': 54 if cross else 108}[variant]
    assert sd['out_net.7.0.weight'].shape == (oc, 140 if cross else 280, 2, 2)
    assert sd['out_net.7.2.weight'].shape == (oc, oc, 2, 2)
    assert m.steps == (54 if cross else 108) and m.disp_min == -3.5 and m.disp_max == 3.5


def test_state_dict_keys_equal_reference_fixture(golden):
    from mmlf_b200.model.feed_forward import FeedForward
    for name, variant, cross, kw in [('net_tiny_upr_full', 'upr', False, {}), ('net_tiny_dpp_cross', 'dpp', True, {}),
                                     ('net_tiny_base_nobn', 'base', False, {'model_no_batchnorm': True})]:
        g = golden(name + '.npz')
        ref = {k[6:]: g[k].shape for k in g.files if k.startswith('state/')}
        m = FeedForward(**fx.model_kwargs(variant, cross, chs=8, **kw))
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert mine == {k: tuple(v) for k, v in ref.items()}


def test_init_matches_reference_rng_stream(golden):
    """Same construction order as the reference => same default init under the same seed (used by the full-width
    fixtures, which store no weights).  The checkpoint written by the reference under seed 5 carries its init."""
    from mmlf_b200.model.feed_forward import FeedForward
    state = torch.load(os.path.join(ROOT, 'tests', 'golden', 'ref_checkpoint_tiny.pt'))
    torch.manual_seed(5)
    m = FeedForward(**state['hyper_parameters'])
    # the reference took one Adam step after init: biases of the last conv moved by at most lr = 1e-3
    for k, v in m.state_dict().items():
        if v.dtype.is_floating_point and 'running' not in k:
            assert (v - state['model_state_dict'][k]).abs().max() <= 1.1e-3, k
    assert set(state.keys()) == {'model_state_dict', 'optimizer_state_dict', 'hyper_parameters', 'epoch', 'iteration',
                                 'loss'}


def test_other_topologies_construct_with_the_reference_layout(golden):
    """--model_ksize 3 and --model_unet: same state_dict keys / shapes as the reference fixtures."""
    from mmlf_b200.engine import Engine
    from mmlf_b200.engine_generic import GenericEngine
    from mmlf_b200.model.feed_forward import FeedForward
    m3 = FeedForward(**fx.model_kwargs('base', chs=8, model_ksize=3))
    g3 = golden('net_tiny_base_k3.npz')
    want = {k[6:]: g3[k].shape for k in g3.files if k.startswith('state/')}
    assert {k: tuple(v.shape) for k, v in m3.state_dict().items()} == want
    assert (m3.padding1, m3.padding2) == (1, 1) and isinstance(m3.engine, GenericEngine)
    mu = FeedForward(**fx.model_kwargs('upr', chs=8, model_unet=True))
    gu = golden('net_unet_upr.npz')
    names = [str(n) for n in gu['names']]
    shapes = [tuple(int(x) for x in s.split(',')) if s else () for s in (str(t) for t in gu['shapes'])]
    assert [(k, tuple(v.shape)) for k, v in mu.state_dict().items()] == list(zip(names, shapes))
    assert isinstance(mu.engine, GenericEngine) and mu.out_chs == 2
    assert isinstance(FeedForward(**fx.model_kwargs('base', chs=8)).engine, Engine)
    with pytest.raises(NotImplementedError):
        FeedForward(**fx.model_kwargs('base', model_ksize=9))
    with pytest.raises(NotImplementedError):
        FeedForward(**fx.model_kwargs('dpp', model_unet=True))


def test_lazy_outputs():
    from mmlf_b200.model.feed_forward import LazyOutputs
    calls = []

    def heavy():
        calls.append(1)
        return {'posterior': 'P', 'one_hot': 'O'}
    o = LazyOutputs({'mean': 1, 'scores': None}, {'posterior': heavy, 'one_hot': heavy})
    assert set(o.keys()) == {'mean', 'scores', 'posterior', 'one_hot'} and not calls
    assert o['mean'] == 1 and not calls
    assert o.get('posterior') == 'P' and o['one_hot'] == 'O' and len(calls) == 1
    assert o.get('missing', 7) == 7
    assert dict(o.items())['posterior'] == 'P'


def test_model_saver_roundtrip(tmp_path):
    from mmlf_b200.model.feed_forward import FeedForward
    from mmlf_b200.utils.dl import ModelSaver
    kw = fx.model_kwargs('upr', False, chs=8)
    m = FeedForward(**kw)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    f = str(tmp_path / 'checkpoint.pt')
    ModelSaver()(f, torch.nn.DataParallel(m), opt, dict(kw, model_radius=11), None, 3, 0.25)
    st = torch.load(f)
    assert set(st) == {'model_state_dict', 'optimizer_state_dict', 'hyper_parameters', 'epoch', 'iteration', 'loss'}
    assert st['iteration'] == 3 and not any(k.startswith('module.') for k in st['model_state_dict'])
    m2 = FeedForward(**st['hyper_parameters'])
    m2.load_state_dict(st['model_state_dict'])


def test_create_mask_margin_and_view_indices(golden):
    from mmlf_b200.data import hci4d
    from mmlf_b200.model import loss
    g = golden('bins.npz')
    for mg in (0, 3, 11):
        assert np.array_equal(loss.create_mask_margin((2, 30, 26), mg).numpy(), g[f'margin{mg}'])
        assert np.array_equal(hci4d.create_mask_margin((2, 30, 26), mg).numpy(), g[f'margin{mg}'])
    gi = golden('indices.npz')
    for n in (9, 7, 5):
        us, vs, ids, dds = hci4d.view_indices((n, n))
        assert us == list(gi[f'us{n}']) and vs == list(gi[f'vs{n}']) and ids == list(gi[f'ids{n}']) and \
            dds == list(gi[f'dds{n}'])
    with pytest.raises(AssertionError):
        hci4d.Shift(1)          # the reference requires a python float (hci4d.py:904)


def test_cli_surfaces_match_the_reference():
    """Every option of the reference's train / validate commands exists here with the same default, flag-ness and type
    (tests/golden/cli_options.json: click introspection of the reference, oracle/gen_golden.py::gen_cli)."""
    import json
    from mmlf_b200.train.cli import main as tmain
    from mmlf_b200.validate.cli import main as vmain
    ref = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'cli_options.json')))
    for name, cmd, extra in (('train', tmain, {'max_iterations', 'gpu_augment', 'synthetic_data'}),
                             ('validate', vmain, {'size', 'synthetic_data'})):
        mine = {p.name: p for p in cmd.params}
        assert set(mine) - {r['name'] for r in ref[name]} == extra
        for r in ref[name]:
            p = mine[r['name']]
            assert list(p.opts) == r['opts'] and type(p).__name__ == r['kind'], r['name']
            assert bool(getattr(p, 'is_flag', False)) == r['is_flag'] and p.type.name == r['type'], r['name']
            assert str(p.default) == str(r['default']), (r['name'], p.default, r['default'])


def test_pfm_roundtrip_and_scene_file_selection(tmp_path):
    """Host side of the HCI4D loader (hci4d.py:129-138, 196-213; utils/pfm.py): file filters, ground-truth pick, PFM I/O."""
    from mmlf_b200.data import hci4d
    from mmlf_b200.utils import dl, pfm
    rng = np.random.RandomState(0)
    img = rng.uniform(-3, 3, (7, 5)).astype(np.float32)
    pfm.save(str(tmp_path / 'a.pfm'), img)
    assert np.array_equal(pfm.load(str(tmp_path / 'a.pfm')), img)
    fx.write_pfm(str(tmp_path / 'b.pfm'), img)                      # independent writer -> our reader
    assert np.array_equal(pfm.load(str(tmp_path / 'b.pfm')), img)
    col = rng.uniform(0, 1, (4, 6, 3)).astype(np.float32)
    pfm.save(str(tmp_path / 'c.pfm'), col)
    assert np.array_equal(pfm.load(str(tmp_path / 'c.pfm')), col)
    with pytest.raises(Exception):
        pfm.save(str(tmp_path / 'd.pfm'), img.astype(np.float64))
    (tmp_path / 'e.pfm').write_bytes(b'P6\n1 1\n-1\n0000')
    with pytest.raises(Exception):
        pfm.load(str(tmp_path / 'e.pfm'))
    files = [f'input_Cam{j:03d}.png' for j in (2, 0, 1)] + ['mask.png', 'a_normals.png', 'objectids.png', 'x_edges.jpg',
                                                             'specular.png', 'unused_1.png', 'gt_disp_lowres.pfm', 'notes.txt']
    assert hci4d.scene_view_files(files) == ['input_Cam000.png', 'input_Cam001.png', 'input_Cam002.png']
    assert hci4d.pick_gt_file(['x.txt'], 40) is None
    assert hci4d.pick_gt_file(['only.pfm'], 40) == 'only.pfm'
    assert hci4d.pick_gt_file(['gt_depth_lowres.pfm', 'gt_disp_lowres.pfm'], 40) == 'gt_disp_lowres.pfm'
    assert hci4d.pick_gt_file(['gt_disp_highres.pfm', 'gt_disp_lowres.pfm'], 40) == 'gt_disp_lowres.pfm'
    assert hci4d.pick_gt_file(['gt_disp_lowres_Cam040.pfm', 'gt_disp_lowres_Cam000.pfm'], 40) == 'gt_disp_lowres_Cam040.pfm'
    # save_img: values outside [0, 1] are min-max normalised, CHW -> HWC, rint(x * 255)
    from PIL import Image
    dl.save_img(str(tmp_path / 'g.png'), np.array([[0.0, 0.5], [1.0, 0.25]], np.float32))
    assert np.asarray(Image.open(str(tmp_path / 'g.png'))).tolist() == [[0, 128], [255, 64]]
    dl.save_img(str(tmp_path / 'h.png'), np.array([[-1.0, 3.0]], np.float32))
    assert np.asarray(Image.open(str(tmp_path / 'h.png'))).tolist() == [[0, 255]]
    dl.save_img(str(tmp_path / 'i.png'), torch.zeros(3, 2, 4))
    assert np.asarray(Image.open(str(tmp_path / 'i.png'))).shape == (2, 4, 3)


def test_hci4d_requires_scene_directories(tmp_path):
    from mmlf_b200.data import hci4d
    with pytest.raises(FileNotFoundError):
        hci4d.HCI4D(str(tmp_path))
    (tmp_path / 'boxes').mkdir()
    ds = hci4d.HCI4D(str(tmp_path), device='cpu')
    assert ds.scenes_names == ['boxes'] and len(ds) == 1
    with pytest.raises(FileNotFoundError):
        ds.load_scene(0)                                            # no view images in the scene


def test_bench_flop_accounting_is_self_consistent():
    """bench.py's per-width split of the conv work (roofline.by_layer_width) adds up to the whole-network figures it is
    reported beside: wide + narrow + head == net_forward_flops; training = forward + data gradients of everything but the
    first conv of each stream; the narrow layers sit below the tensor/HBM ridge (they are HBM-bound), the wide ones above."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    B, H, W = 64, 96, 96
    for variant, oc in (('base', 1), ('upr', 2), ('dpp', 108)):
        fwd = bench.net_forward_flops(B, H, W, variant)
        wide, narrow, nbytes = bench.conv_split(B, H, W, variant, False)
        head = bench.conv_flops(B, H, W, 280, oc, 0) + bench.conv_flops(B, H, W, oc, oc, 1)
        assert abs(wide + narrow + head - fwd) <= 1e-9 * fwd
        wide_t, narrow_t, nbytes_t = bench.conv_split(B, H, W, variant, True)
        first = 4 * bench.conv_flops(B, H, W, 27, 70, 0)
        assert abs(wide_t - 2 * wide) <= 1e-9 * wide and abs(narrow_t - (2 * narrow - first)) <= 1e-9 * narrow
        assert nbytes_t > 2 * nbytes                     # twins, ReLU bits and the data gradients' traffic
        # arithmetic intensity against the ridge of the measured peaks (fallback: 1400 TFLOP/s / 6.5 TB/s)
        try:
            pk = json.load(open(os.path.join(root, 'MEASURED_PEAKS.json')))
            ridge = pk['bf16_tflops_sustained'] * 1e12 / (pk['hbm_gbs'] * 1e9)
        except (OSError, KeyError):
            ridge = 1400e12 / 6500e9
        assert narrow / nbytes < ridge and narrow_t / nbytes_t < ridge
        n_slots = B * (H + 1) * (W + 1)
        assert bench.conv_flops(B, H, W, 280, 280, 1) / (n_slots * 2.0 * 2 * 288) > ridge


def test_profiles_readme_is_regenerable():
    """profiles/README.md is written by tools/profile_tables.py from the evidence files next to it: the committed text is
    what the tool prints today (so the numbers in it are the files' numbers, not typed ones)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'profile_tables.py'), '--readme', 'r02'],
                       capture_output=True, text=True, cwd=root, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    committed = open(os.path.join(root, 'profiles', 'README.md')).read()
    # MEASURED_PEAKS.json is driver-written and may be absent or differ on another box: compare everything below the header
    cut = '## Test / parity evidence'
    assert cut in r.stdout and cut in committed
    assert r.stdout[r.stdout.index(cut):].strip() == committed[committed.index(cut):].strip()
