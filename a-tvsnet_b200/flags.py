"""Module-level configuration the reference ops read from ``tf.app.flags.FLAGS``
(example.py:27-48; read inside ops at homography_warping.py:149,215,301,369,
model.py:96,248).  Same field names and defaults; ``precision`` is new."""


class _Flags(object):
    def __init__(self):
        self.view_num = 5            # example.py:32
        self.max_d = 128             # example.py:37
        self.sample_scale = 0.25     # example.py:43
        self.batch_size = 1          # example.py:45
        self.inverse_depth = True    # example.py:47
        self.num_gpus = 1
        self.gpu_id = 0
        # 'fp32' = CUDA-core parity path, 'bf16' = tcgen05 tensor-core path for the 3-D CNN
        self.precision = 'bf16'
        # tensor-core path: dtype of the raw (pre-BN) convolution outputs, 'f16' (saturated) or 'f32'
        self.raw_dtype = 'f16'


FLAGS = _Flags()
