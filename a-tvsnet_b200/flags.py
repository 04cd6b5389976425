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
        # 2-D feature extractor: run its stride-1 convolutions on the tensor cores (fp16 operands, fp32 residual stream
        # and statistics; needs precision == 'fp16').  OFF by default: 5 x faster FEM (7.0 -> 2.1 ms per 640x512 image)
        # but ~50 layers of 11-bit operand rounding put 0.6 % of error on the features, which moves the depth map by
        # 0.09 - 0.24 % of the range (tools/fem_accuracy.py) - outside the 0.1 % bound the fp32 FEM keeps (0.015 - 0.03 %)
        self.fem_tensor = False
        # ... and of the layers fed by the UN-normalised cost volume (conv_b0_0_1 / conv_b0_1_0), whose magnitude
        # follows the checkpoint's feature scale: fp16 as well, guarded by the epilogues' saturation counter
        # (atvs_saturation_count, pipeline.check_saturation): a clamped value is detected, never silent; set 'f32'
        # for a checkpoint whose features exceed ~1e3
        self.first_raw_dtype = 'f16'
        # attention aggregation (AAM) of the 16-bit path as ONE kernel (atvs_attention_fused: the attention convolutions of
        # all views + softmax over views + weighted sum, logits kept in TMEM) instead of one 8 -> 16 convolution per view
        # + atvs_attention_raw; the sharded (multi-GPU) aggregation always uses the latter
        self.attention_fused = True


FLAGS = _Flags()
