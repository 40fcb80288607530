"""Independent construction of the original-GTCRN-style architecture (BASELINE configs[3]) from STOCK torch layers --
nn.Unfold, nn.Conv2d / nn.ConvTranspose2d with the upstream padding / dilation / groups arguments, nn.GRU, nn.LayerNorm --
used only to cross-check oracle/gtcrn_full_oracle.py's functional restatement (causal-conv form of the dilated depthwise
ConvTranspose2d, explicit GRU cell loops, manual unfold).  Architecture source: the commented-out lines of
/root/reference/gtcrn_micro/models/gtcrn_micro.py:187,199,346,361,381,391,417-418,429,448,494 plus SURVEY.md Appendix C.
Parity for this variant is unpinned (no reference implementation exists); this file is test infrastructure."""
import torch
import torch.nn as nn


class SFE(nn.Module):
    def __init__(self):
        super().__init__()
        self.unfold = nn.Unfold(kernel_size=(1, 3), stride=(1, 1), padding=(0, 1))

    def forward(self, x):
        return self.unfold(x).reshape(x.shape[0], x.shape[1] * 3, x.shape[2], x.shape[3])


class TRA(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.att_gru = nn.GRU(c, 2 * c, 1, batch_first=True)
        self.att_fc = nn.Linear(2 * c, c)

    def forward(self, x):
        zt = x.pow(2).mean(dim=-1)
        at = torch.sigmoid(self.att_fc(self.att_gru(zt.transpose(1, 2))[0])).transpose(1, 2)
        return x * at[..., None]


class ConvBlock(nn.Module):
    def __init__(self, cin, cout, groups=1, deconv=False, last=False):
        super().__init__()
        mod = nn.ConvTranspose2d if deconv else nn.Conv2d
        self.conv = mod(cin, cout, (1, 5), stride=(1, 2), padding=(0, 2), groups=groups)
        self.bn = nn.BatchNorm2d(cout)
        self.act = nn.Tanh() if last else nn.PReLU()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class GTConvBlock(nn.Module):
    def __init__(self, d, deconv):
        super().__init__()
        mod = nn.ConvTranspose2d if deconv else nn.Conv2d
        self.pad = 2 * d
        self.sfe = SFE()
        self.point_conv1 = mod(24, 16, 1)
        self.point_bn1 = nn.BatchNorm2d(16)
        self.point_act = nn.PReLU()
        self.depth_conv = mod(16, 16, (3, 3), stride=(1, 1), padding=(2 * d, 1) if deconv else (0, 1), dilation=(d, 1), groups=16)
        self.depth_bn = nn.BatchNorm2d(16)
        self.depth_act = nn.PReLU()
        self.point_conv2 = mod(16, 8, 1)
        self.point_bn2 = nn.BatchNorm2d(8)
        self.tra = TRA(8)

    def forward(self, x):
        x1, x2 = torch.chunk(x, 2, dim=1)
        h = self.point_act(self.point_bn1(self.point_conv1(self.sfe(x1))))
        h = nn.functional.pad(h, [0, 0, self.pad, 0])
        h = self.depth_act(self.depth_bn(self.depth_conv(h)))
        h = self.tra(self.point_bn2(self.point_conv2(h)))
        return torch.stack([h, x2], dim=2).reshape(x.shape)


class GRNN(nn.Module):
    def __init__(self, n_in, n_h, bidirectional):
        super().__init__()
        self.rnn1 = nn.GRU(n_in // 2, n_h // 2, 1, batch_first=True, bidirectional=bidirectional)
        self.rnn2 = nn.GRU(n_in // 2, n_h // 2, 1, batch_first=True, bidirectional=bidirectional)

    def forward(self, x):
        x1, x2 = torch.chunk(x, 2, dim=-1)
        return torch.cat([self.rnn1(x1)[0], self.rnn2(x2)[0]], dim=-1)


class DPGRNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.intra_rnn = GRNN(16, 8, True)
        self.intra_fc = nn.Linear(16, 16)
        self.intra_ln = nn.LayerNorm((33, 16), eps=1e-8)
        self.inter_rnn = GRNN(16, 16, False)
        self.inter_fc = nn.Linear(16, 16)
        self.inter_ln = nn.LayerNorm((33, 16), eps=1e-8)

    def forward(self, x):
        B, C, T, Fq = x.shape
        x = x.permute(0, 2, 3, 1)
        a = self.intra_fc(self.intra_rnn(x.reshape(B * T, Fq, C))).reshape(B, T, Fq, C)
        intra = x + self.intra_ln(a)
        e = intra.permute(0, 2, 1, 3).reshape(B * Fq, T, C)
        e = self.inter_fc(self.inter_rnn(e)).reshape(B, Fq, T, C).permute(0, 2, 1, 3)
        return (intra + self.inter_ln(e)).permute(0, 3, 1, 2)


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.en_convs = nn.ModuleList([ConvBlock(9, 16), ConvBlock(16, 16, groups=2), GTConvBlock(1, False), GTConvBlock(2, False), GTConvBlock(5, False)])


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.de_convs = nn.ModuleList([GTConvBlock(5, True), GTConvBlock(2, True), GTConvBlock(1, True),
                                       ConvBlock(16, 16, groups=2, deconv=True), ConvBlock(16, 2, deconv=True, last=True)])


class ERB(nn.Module):
    def __init__(self):
        super().__init__()
        self.erb_fc = nn.Linear(192, 64, bias=False)
        self.ierb_fc = nn.Linear(64, 192, bias=False)


class GTCRNFullModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.erb = ERB()
        self.sfe = SFE()
        self.encoder = Encoder()
        self.dpgrnn1 = DPGRNN()
        self.dpgrnn2 = DPGRNN()
        self.decoder = Decoder()

    def forward(self, spec):
        re, im = spec[..., 0].permute(0, 2, 1), spec[..., 1].permute(0, 2, 1)
        feat = torch.stack([torch.sqrt(re ** 2 + im ** 2 + 1e-12), re, im], dim=1)
        feat = torch.cat([feat[..., :65], self.erb.erb_fc(feat[..., 65:])], dim=-1)
        x = self.sfe(feat)
        outs = []
        for blk in self.encoder.en_convs:
            x = blk(x)
            outs.append(x)
        x = self.dpgrnn2(self.dpgrnn1(x))
        for i, blk in enumerate(self.decoder.de_convs):
            x = blk(x + outs[4 - i])
        m = torch.cat([x[..., :65], self.erb.ierb_fc(x[..., 65:])], dim=-1)
        return torch.stack([re * m[:, 0] - im * m[:, 1], im * m[:, 0] + re * m[:, 1]], dim=-1).permute(0, 2, 1, 3)
