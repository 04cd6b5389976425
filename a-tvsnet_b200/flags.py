"""Module-level configuration the reference ops read from ``tf.app.flags.FLAGS``
(example.py:27-48; read inside ops at homography_warping.py:149,215,301,369,
model.py:96,248).  Same field names and defaults; ``precision`` is new."""


DEFAULT_PRECISION = 'fp16'


class _Flags(object):
    def __init__(self):
        self.view_num = 5            # example.py:32
        self.max_d = 128             # example.py:37
        self.sample_scale = 0.25     # example.py:43
        self.batch_size = 1          # example.py:45
        self.inverse_depth = True    # example.py:47
        self.num_gpus = 1
        self.gpu_id = 0
        # 'fp32' = CUDA-core parity path; 'fp16' | 'bf16' = tcgen05 tensor-core path for the 3-D CNN with
        # activations and weights stored in that 16-bit format (fp32 accumulation, fp32 BN statistics).
        # fp16 is the default: same tensor rate, 11 instead of 8 significant bits on BN-normalised (bounded)
        # activations -> depth MAE 0.02-0.03 % of the range against 0.17-0.19 % for bf16
        self.precision = DEFAULT_PRECISION
        # tensor-core path: dtype of the raw (pre-BN) convolution outputs, 'f16' (saturated) or 'f32'
        self.raw_dtype = 'f16'
        # ... and of the layers fed by the UN-normalised cost volume (conv_b0_0_1 / conv_b0_1_0), whose magnitude
        # follows the checkpoint's feature scale: fp16 as well, guarded by the epilogues' saturation counter
        # (atvs_saturation_count, pipeline.check_saturation): a clamped value is detected, never silent; set 'f32'
        # for a checkpoint whose features exceed ~1e3
        self.first_raw_dtype = 'f16'


FLAGS = _Flags()
