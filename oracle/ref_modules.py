"""TEST INFRASTRUCTURE / CPU arm of bench.py only.  Loader for the reference's OWN DenseFusion Python modules
(lib/network.py, lib/pspnet.py, lib/extractors.py, lib/transformations.py, tools/utils.py), compiled unmodified by Cython
into extension modules under oracle/_ref/DenseFusion (recipe: oracle/Makefile, built in the dev container where
/root/reference exists; the .so files travel to the GPU box, the sources do not).  Nothing in the product imports this.
"""
import os
import sys
import warnings

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return all(os.path.exists(os.path.join(_REF, 'DenseFusion', p)) for p in
               ('lib/network.so', 'lib/pspnet.so', 'lib/extractors.so', 'lib/transformations.so', 'tools/utils.so'))


def load():
    """-> (network, tools_utils, transformations) modules of the compiled reference; raises ImportError if absent."""
    if not available():
        raise ImportError('oracle/_ref/DenseFusion is not built (run `make -C oracle ref` where /root/reference exists)')
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    warnings.filterwarnings('ignore')
    import DenseFusion.lib.network as network
    import DenseFusion.tools.utils as tools
    import DenseFusion.lib.transformations as tf
    for m in (network, tools, tf):
        assert os.path.abspath(m.__file__).startswith(_REF), 'not the compiled reference: ' + m.__file__
    return network, tools, tf


def canonical_prediction(mods, estimator, refiner, out_img, cloud, choose, idx, num_points, iterations=2):
    """One object through the reference's own modules: PoseNet (cnn = Identity: `out_img` is the encoder output) ->
    my_estimator_prediction -> `iterations` x (cloud re-expressed in the current pose, PoseRefineNet,
    my_refined_prediction).  The loop glue follows DenseFusion/tools/eval_linemod.py:81-114 (script code there, not a
    function of the reference); every arithmetic step is a call into the compiled reference modules.
    Returns (q [4] wxyz, t [3])."""
    import numpy as np
    import torch
    network, tools, tf = mods
    pred_r, pred_t, pred_c, emb = estimator(out_img, cloud, choose, idx)
    _, my_r, my_t = tools.my_estimator_prediction(pred_r, pred_t, pred_c, num_points, 1, cloud)
    for _ in range(iterations):
        T = torch.from_numpy(my_t.astype(np.float32)).view(1, 3).repeat(num_points, 1).contiguous().view(1, num_points, 3)
        my_mat = tf.quaternion_matrix(my_r)
        R = torch.from_numpy(my_mat[:3, :3].astype(np.float32)).view(1, 3, 3)
        new_cloud = torch.bmm((cloud - T), R).contiguous()
        pred_r, pred_t = refiner(new_cloud, emb, idx)
        _, my_r, my_t = tools.my_refined_prediction(pred_r, pred_t, my_r, my_t)
    return my_r, my_t
