"""``mmlf.utils.lf.save_views`` (/root/reference/mmlf/utils/lf.py:6-55): dump the view stacks of a scene as PNG files
`view_{h,v,i,d}_{j}.png` (host-side file I/O used by ``HCI4D.save_batch``)."""
import os

from . import dl


def save_views(scene_dir, h_views, v_views, i_views=None, d_views=None):
    os.makedirs(scene_dir, exist_ok=True)
    for tag, stack in (('h', h_views), ('v', v_views), ('i', i_views), ('d', d_views)):
        if stack is None:
            continue
        if len(stack.shape) == 5:                    # drop a batch dimension (lf.py:25-30)
            stack = stack[0]
        for j in range(stack.shape[0]):
            dl.save_img(os.path.join(scene_dir, f'view_{tag}_{j}.png'), stack[j])
