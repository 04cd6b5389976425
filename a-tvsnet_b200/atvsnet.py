"""Network definitions of /root/reference/cnn_wrapper/atvsnet.py for the hot path:
StackedUNet :5-96, StackedUNet_prob :100-192 (CRM), AttAggregation(_keepchannel) :196-213,
OutputConv(_refine) :216-226, AttAggregation_refine(_keepchannel) :229-242 - same class
names, layer names (= checkpoint variable scopes) and ``Cls({'data': t}, is_training=True,
reuse=...)`` construction.  The three hourglass blocks share one body here instead of being
written out three times."""
from .network import Network


class StackedUNet_prob(Network):
    with_prob = True

    def _block(self, b):
        bf = 8
        p = 'conv_b%d' % b
        src = 'data' if b == 0 else p + '_0_0'
        (self.feed(src)
             .conv_bn(3, bf * 2, 2, name=p + '_1_0')
             .conv_bn(3, bf * 4, 2, name=p + '_2_0')
             .conv_bn(3, bf * 8, 2, name=p + '_3_0'))
        self.feed(src).conv_bn(3, bf, 1, name=p + '_0_1')
        if b == 0:
            self.feed(p + '_1_0').conv_bn(3, bf * 2, 1, name=p + '_1_1')
            self.feed(p + '_2_0').conv_bn(3, bf * 4, 1, name=p + '_2_1')
            long1, long2 = (), ()
        else:
            q = 'conv_b%d' % (b - 1)
            (self.feed(p + '_1_0', q + '_5_0').add(name=p + '_1_1_concat')
                 .conv_bn(3, bf * 2, 1, name=p + '_1_1'))
            (self.feed(p + '_2_0', q + '_4_0').add(name=p + '_2_1_concat')
                 .conv_bn(3, bf * 4, 1, name=p + '_2_1'))
            long1, long2 = ('conv_b0_1_1',), ('conv_b0_2_1',)     # skips to block 0 (:152-157, :182-187)
        (self.feed(p + '_3_0')
             .conv_bn(3, bf * 8, 1, name=p + '_3_1')
             .deconv_bn(3, bf * 4, 2, name=p + '_4_0'))
        (self.feed(p + '_4_0', p + '_2_1', *long2).add(name=p + '_4_1')
             .deconv_bn(3, bf * 2, 2, name=p + '_5_0'))
        (self.feed(p + '_5_0', p + '_1_1', *long1).add(name=p + '_5_1')
             .deconv_bn(3, bf, 2, name=p + '_6_0'))
        joined = 'conv_b%d_0_0' % (b + 1) if b < 2 else 'conv_b2_6_1'
        self.feed(p + '_6_0', p + '_0_1').add(name=joined)

    def setup(self):
        shp = tuple(self.inputs['data'].shape)
        if len(shp) != 5 or any(int(n) % 8 for n in shp[1:4]):
            # three stride-2 convolutions, then three stride-2 transposed convolutions joined by `add`
            # (atvsnet.py:104-128): TF raises a shape error at the first join otherwise
            raise ValueError("StackedUNet: input must be (B,D,h,w,C) with D, h, w multiples of 8, got %s" % (shp,))
        for b in range(3):
            self._block(b)
        if self.with_prob:
            self.feed('conv_b2_6_1').conv(3, 1, 1, relu=False, name='conv_b2_6_2')


class StackedUNet(StackedUNet_prob):
    with_prob = False


class AttAggregation_keepchannel(Network):
    scope = 'attention_aggregate'

    def setup(self):
        # data size of (B, D, H, W, C, N), N=NumNeigh
        (self.feed('data')
             .attention_aggregation(kernel_size=3, name=self.scope, second_weight=True, relu=True, biased=False))


class AttAggregation(Network):
    scope, prob = 'attention_aggregate', 'attention_prob_vol'

    def setup(self):
        (self.feed('data')
             .attention_aggregation(kernel_size=3, name=self.scope, second_weight=True, relu=True, biased=False)
             .conv(3, 1, 1, relu=False, name=self.prob))


class AttAggregation_refine_keepchannel(AttAggregation_keepchannel):
    scope = 'attention_aggregate_refine'


class AttAggregation_refine(AttAggregation):
    scope, prob = 'attention_aggregate_refine', 'attention_prob_vol_refine'


class OutputConv(Network):
    prob = 'attention_prob_vol'

    def setup(self):
        self.feed('data').conv(3, 1, 1, relu=False, name=self.prob)


class OutputConv_refine(OutputConv):
    prob = 'attention_prob_vol_refine'
