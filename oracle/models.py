"""Oracle restatement of the reference generator / discriminator families (CPU, fp32).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference builds its networks inside five copy-pasted scripts that read a
module-level ``PATCH_SHAPE`` global.  Here the same networks are assembled from small
layer specs with the patch size passed explicitly; parameter names, shapes and
registration order are kept identical so that a reference ``state_dict`` loads with
``strict=True`` (checked in tests/test_oracle_golden.py).

Reference anchors:
  base model      vae-gan.py:47-84 (encoder/decoder), :118-146 (VAEGAN), :148-159 (D)
  U-Net + FiLM    vae-gan-v2.py:65-114 (char encoder), :117-149 (FiLM), :152-187 (encoder),
                  :191-280 (decoder), :283-327 (VAEGAN_UNet_SpatialFiLM)
  U-Net (no FiLM) vae-gan-unet.py:124-176 (encoder), :179-254 (decoder), :257-297 (model)
"""
from __future__ import annotations

import zlib
from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

ALPHABET_STR = (" !\"#$%&'()*+,-./0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[\\]^_`"
                "abcdefghijklmnopqrstuvwxyz{|}~")  # vae-gan-v2.py:33
# vae-gan-unet.py:33 extends the alphabet with the Russian letters (incl. yo) in both cases
_RU_LOWER = "".join(chr(c) for c in range(0x430, 0x436)) + "\u0451" + "".join(chr(c) for c in range(0x436, 0x450))
_RU_UPPER = "".join(chr(c) for c in range(0x410, 0x416)) + "\u0401" + "".join(chr(c) for c in range(0x416, 0x430))
ALPHABET_STR_UNET = ALPHABET_STR + _RU_LOWER + _RU_UPPER
SBERT_DIM = 384  # paraphrase-multilingual-MiniLM-L12-v2 sentence embedding width (vae-gan.py:32)


# --------------------------------------------------------------------------------------
# small builders
# --------------------------------------------------------------------------------------
def _bn_relu(ch: int) -> List[nn.Module]:
    return [nn.BatchNorm2d(ch), nn.ReLU(inplace=True)]


def double_conv(cin: int, cout: int) -> nn.Sequential:
    """[Conv3x3 p1 (no bias) -> BN -> ReLU] x2   (vae-gan-v2.py:171-177, :236-242)."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), *_bn_relu(cout),
                         nn.Conv2d(cout, cout, 3, padding=1, bias=False), *_bn_relu(cout))


def hash_sentence_embedding(texts: Sequence[str], dim: int = SBERT_DIM) -> torch.Tensor:
    """Deterministic stand-in for ``SentenceTransformer.encode`` (vae-gan.py:110).

    SBERT weights cannot be fetched offline, so both the reference (through a stub class,
    see tests/golden/make_golden.py) and this oracle embed a string as a unit-variance vector seeded by
    the CRC32 of its UTF-8 bytes.  Parity of the *real* text model is unpinned.
    """
    rows = []
    for t in texts:
        g = torch.Generator().manual_seed(zlib.crc32(t.encode("utf-8")))
        rows.append(torch.randn(dim, generator=g))
    return torch.stack(rows)


# --------------------------------------------------------------------------------------
# base conv VAE-GAN  (vae-gan.py)
# --------------------------------------------------------------------------------------
class VAEEncoder(nn.Module):
    """vae-gan.py:47-66 -- four stride-2 3x3 convs (+bias) with BN/ReLU, two full-kernel heads."""

    WIDTHS = (128, 256, 512, 1024)

    def __init__(self, in_ch: int, z_ch: int, patch_hw: Tuple[int, int]):
        super().__init__()
        layers, c = [], in_ch
        for w in self.WIDTHS:
            layers += [nn.Conv2d(c, w, 3, 2, 1), *_bn_relu(w)]
            c = w
        self.feat = nn.Sequential(*layers)
        k = (patch_hw[0] // 16, patch_hw[1] // 16)
        self.mu_head = nn.Conv2d(c, z_ch, kernel_size=k)
        self.logvar_head = nn.Conv2d(c, z_ch, kernel_size=k)

    def forward(self, x):
        h = self.feat(x)
        return self.mu_head(h), self.logvar_head(h)


class VAEDecoder(nn.Module):
    """vae-gan.py:68-84 -- full-kernel ConvT from 1x1, four ConvT4x4 s2 p1, Conv3x3, Sigmoid."""

    WIDTHS = (1024, 512, 256, 128, 64)

    def __init__(self, z_ch: int, text_ch: int, out_ch: int, patch_hw: Tuple[int, int]):
        super().__init__()
        k = (patch_hw[0] // 16, patch_hw[1] // 16)
        w = self.WIDTHS
        layers: List[nn.Module] = [nn.ConvTranspose2d(z_ch + text_ch, w[0], kernel_size=k), *_bn_relu(w[0])]
        for a, b in zip(w[:-1], w[1:]):
            layers += [nn.ConvTranspose2d(a, b, 4, 2, 1), *_bn_relu(b)]
        layers += [nn.Conv2d(w[-1], out_ch, 3, 1, 1), nn.Sigmoid()]
        self.decode = nn.Sequential(*layers)

    def forward(self, zc):
        return self.decode(zc)


class SentenceTextEncoder(nn.Module):
    """vae-gan.py:86-116 with the SBERT model replaced by an injected ``embedder`` callable.

    Only ``fc`` (Linear 384 -> text_ch) is trainable in the reference as well: ``encode`` runs
    without grad (vae-gan.py:110).
    """

    def __init__(self, out_dim: int, embedder: Callable[[Sequence[str]], torch.Tensor] = hash_sentence_embedding):
        super().__init__()
        self.embedder = embedder
        self.fc = nn.Linear(SBERT_DIM, out_dim)

    def forward(self, texts):
        with torch.no_grad():
            e = self.embedder(texts).to(self.fc.weight)
        return self.fc(e)


def reparameterize(mu, logvar):
    """vae-gan.py:133-136 (same at vae-gan-v2.py:312-315, vae-gan-unet.py:283-286)."""
    std = torch.exp(0.5 * logvar)
    return mu + torch.randn_like(std) * std


class VAEGAN(nn.Module):
    """vae-gan.py:124-146."""

    def __init__(self, in_ch=4, z_ch=128, text_ch=64, out_ch=3, patch_hw=(64, 448)):
        super().__init__()
        self.encoder = VAEEncoder(in_ch, z_ch, patch_hw)
        self.text_encoder = SentenceTextEncoder(text_ch)
        self.decoder = VAEDecoder(z_ch, text_ch, out_ch, patch_hw)

    def forward(self, image, mask, texts):
        mu, logvar = self.encoder(torch.cat([image, mask], 1))
        z = reparameterize(mu, logvar)
        t = self.text_encoder(texts)
        t = t[:, :, None, None].expand(-1, -1, z.shape[2], z.shape[3])  # spatial_broadcast, :118-122
        return self.decoder(torch.cat([z, t], 1)), mu, logvar


class Discriminator(nn.Module):
    """vae-gan.py:148-159 (byte-equivalent at vae-gan-v2.py:330-349, vae-gan-unet.py:299-317)."""

    def __init__(self, in_ch=3):
        super().__init__()
        body: List[nn.Module] = []
        c = in_ch
        for i, w in enumerate((64, 128, 256, 512)):
            body.append(spectral_norm(nn.Conv2d(c, w, 4, 2, 1)))
            if i:
                body.append(nn.InstanceNorm2d(w, affine=True))
            body.append(nn.LeakyReLU(0.2, inplace=True))
            c = w
        body.append(nn.Conv2d(c, 1, 4, 1, 1))
        self.body = nn.Sequential(*body)

    def forward(self, x):
        return self.body(x)


# --------------------------------------------------------------------------------------
# U-Net families  (vae-gan-v2.py, vae-gan-unet.py)
# --------------------------------------------------------------------------------------
class CharacterTokenEncoder(nn.Module):
    """vae-gan-v2.py:65-114 -- char vocab -> Embedding -> biGRU -> adaptive avg pool to W/16."""

    def __init__(self, alphabet_str, emb_dim, rnn_hidden_dim, rnn_layers, target_feature_width):
        super().__init__()
        self.lut = {ch: i + 1 for i, ch in enumerate(alphabet_str)}
        self.embedding = nn.Embedding(len(alphabet_str) + 1, emb_dim, padding_idx=0)
        self.rnn = nn.GRU(emb_dim, rnn_hidden_dim, num_layers=rnn_layers, batch_first=True,
                          bidirectional=True, dropout=0.1 if rnn_layers > 1 else 0)
        self.rnn_output_dim = 2 * rnn_hidden_dim
        self.adaptive_pool = nn.AdaptiveAvgPool1d(target_feature_width)

    def tokenize(self, texts, max_len=60):
        idx = torch.zeros(len(texts), max_len, dtype=torch.long)
        for r, t in enumerate(texts):
            ids = [self.lut.get(ch, 0) for ch in t][:max_len]
            idx[r, :len(ids)] = torch.tensor(ids, dtype=torch.long)
        return idx

    def forward(self, texts, max_len=60):
        idx = self.tokenize(texts, max_len).to(self.embedding.weight.device)
        out, _ = self.rnn(self.embedding(idx))
        return self.adaptive_pool(out.permute(0, 2, 1)).unsqueeze(2)  # (B, 2*hid, 1, W/16)


class VAEEncoderWithSkips(nn.Module):
    """vae-gan-v2.py:152-187 == vae-gan-unet.py:124-176."""

    def __init__(self, in_ch, z_ch, patch_hw):
        super().__init__()
        self.e_conv1 = double_conv(in_ch, 64)
        self.pool1 = nn.MaxPool2d(2, 2)
        self.e_conv2 = double_conv(64, 128)
        self.pool2 = nn.MaxPool2d(2, 2)
        self.e_conv3 = double_conv(128, 256)
        self.pool3 = nn.MaxPool2d(2, 2)
        self.e_conv4 = double_conv(256, 512)
        self.pool4 = nn.MaxPool2d(2, 2)
        self.bottleneck_conv = double_conv(512, 1024)
        k = (patch_hw[0] // 16, patch_hw[1] // 16)
        self.mu_head = nn.Conv2d(1024, z_ch, kernel_size=k)
        self.logvar_head = nn.Conv2d(1024, z_ch, kernel_size=k)

    def forward(self, x, return_pooled=False):
        skips, pooled = [], []
        for i in (1, 2, 3, 4):
            x = getattr(self, f"e_conv{i}")(x)
            skips.append(x)
            x = getattr(self, f"pool{i}")(x)
            pooled.append(x)
        b = self.bottleneck_conv(x)
        out = (self.mu_head(b), self.logvar_head(b), skips)
        return out + (pooled,) if return_pooled else out


class SpatialFiLMLayer(nn.Module):
    """vae-gan-v2.py:117-149."""

    def __init__(self, text_channels_in, num_features_main):
        super().__init__()
        t = text_channels_in
        self.param_predictor = nn.Sequential(nn.Conv2d(t, t, 3, padding=1, bias=False), *_bn_relu(t),
                                             nn.Conv2d(t, 2 * num_features_main, 1))
        self.num_features_main = num_features_main

    def forward(self, x_main, text_base):
        t = F.interpolate(text_base, size=x_main.shape[2:], mode="bilinear", align_corners=False)
        gb = self.param_predictor(t)
        n = self.num_features_main
        return gb[:, :n] * x_main + gb[:, n:]


class VAEDecoderWithSpatialFiLM(nn.Module):
    """vae-gan-v2.py:191-280."""

    def __init__(self, z_ch, text_channels_in, out_ch_image, patch_h, patch_w):
        super().__init__()
        self.initial_h, self.initial_w = patch_h // 16, patch_w // 16
        self.bottleneck_proc = nn.Sequential(
            nn.ConvTranspose2d(z_ch + text_channels_in, 1024, kernel_size=(self.initial_h, 1)), *_bn_relu(1024))
        c = 1024
        for i, skip in enumerate((512, 256, 128, 64), start=1):
            setattr(self, f"up_tconv{i}", nn.ConvTranspose2d(c, c // 2, 2, 2))
            cat = c // 2 + skip
            setattr(self, f"spatial_film{i}", SpatialFiLMLayer(text_channels_in, cat))
            setattr(self, f"conv_block{i}", double_conv(cat, c // 2))
            c //= 2
        self.final_image_conv = nn.Conv2d(c, out_ch_image, 1)

    def forward(self, z, text_base, skips):
        x = self.bottleneck_proc(torch.cat([z.expand(-1, -1, 1, self.initial_w), text_base], 1))
        for i in (1, 2, 3, 4):
            x = torch.cat([getattr(self, f"up_tconv{i}")(x), skips[4 - i]], 1)
            x = getattr(self, f"spatial_film{i}")(x, text_base)
            x = getattr(self, f"conv_block{i}")(x)
        return torch.sigmoid(self.final_image_conv(x))


class VAEGAN_UNet_SpatialFiLM(nn.Module):
    """vae-gan-v2.py:283-327."""

    def __init__(self, in_ch_style=4, z_ch_style=128, out_ch_img=3, alphabet_str_text=ALPHABET_STR,
                 char_emb_dim_text=128, char_rnn_hidden_dim_text=256, char_rnn_layers_text=2,
                 patch_hw=(64, 448)):
        super().__init__()
        self.char_text_encoder_module = CharacterTokenEncoder(
            alphabet_str_text, char_emb_dim_text, char_rnn_hidden_dim_text, char_rnn_layers_text, patch_hw[1] // 16)
        self.style_vae_encoder_module = VAEEncoderWithSkips(in_ch_style, z_ch_style, patch_hw)
        self.image_vae_decoder_module = VAEDecoderWithSpatialFiLM(
            z_ch_style, self.char_text_encoder_module.rnn_output_dim, out_ch_img, patch_hw[0], patch_hw[1])

    def forward(self, image, mask, texts):
        mu, logvar, skips = self.style_vae_encoder_module(torch.cat([image, mask], 1))
        z = reparameterize(mu, logvar)
        text_base = self.char_text_encoder_module(texts)
        return self.image_vae_decoder_module(z, text_base, skips), mu, logvar


def up_block(cin: int, cout: int) -> nn.Sequential:
    """vae-gan-unet.py:210-222 -- ConvT2x2 s2 -> BN -> ReLU -> 2x(Conv3x3 -> BN -> ReLU)."""
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 2, 2), *_bn_relu(cout),
                         nn.Conv2d(cout, cout, 3, padding=1, bias=False), *_bn_relu(cout),
                         nn.Conv2d(cout, cout, 3, padding=1, bias=False), *_bn_relu(cout))


class VAEDecoderWithSkips(nn.Module):
    """vae-gan-unet.py:179-254.

    ``forward`` follows the reference line by line and therefore raises the same shape error the
    reference raises.  ``forward_repaired`` is the documented minimal
    repair that keeps every parameter shape: the bottleneck ConvT is fed a 1x1 input (text map
    mean-pooled over width) and the *pooled* encoder maps are used as skips.
    """

    def __init__(self, z_ch, text_feat_channels, out_ch_image, patch_hw):
        super().__init__()
        self.initial_h, self.initial_w = patch_hw[0] // 16, patch_hw[1] // 16
        self.bottleneck_upsample = nn.Sequential(
            nn.ConvTranspose2d(z_ch + text_feat_channels, 1024, kernel_size=(self.initial_h, self.initial_w)),
            *_bn_relu(1024))
        c = 1024
        for i, skip in enumerate((512, 256, 128, 64), start=1):
            setattr(self, f"d_upconv{i}", up_block(c + skip, c // 2))
            c //= 2
        self.final_image_conv = nn.Conv2d(c, out_ch_image, 1)

    def forward(self, z, text_feat, skips):
        x = self.bottleneck_upsample(torch.cat([z.expand(-1, -1, 1, self.initial_w), text_feat], 1))
        for i in (1, 2, 3, 4):
            x = getattr(self, f"d_upconv{i}")(torch.cat([x, skips[4 - i]], 1))
        return torch.sigmoid(self.final_image_conv(x))

    def forward_repaired(self, z, text_feat, pooled):
        x = self.bottleneck_upsample(torch.cat([z, text_feat.mean(dim=3, keepdim=True)], 1))
        for i in (1, 2, 3, 4):
            x = getattr(self, f"d_upconv{i}")(torch.cat([x, pooled[4 - i]], 1))
        return torch.sigmoid(self.final_image_conv(x))


class VAEGAN_UNet_CharEmb(nn.Module):
    """vae-gan-unet.py:257-297; ``repaired=True`` selects the row-U repair (the default here,
    because the unrepaired forward cannot run for any PATCH_SHAPE)."""

    def __init__(self, in_ch_for_style_encoder=4, z_ch_for_style=128, out_ch_for_image=3,
                 alphabet_str_for_text=ALPHABET_STR_UNET, char_emb_dim_for_text=128,
                 char_rnn_hidden_dim_for_text=256, char_rnn_layers_for_text=2, patch_hw=(64, 448), repaired=True):
        super().__init__()
        self.repaired = repaired
        self.char_text_encoder_module = CharacterTokenEncoder(
            alphabet_str_for_text, char_emb_dim_for_text, char_rnn_hidden_dim_for_text,
            char_rnn_layers_for_text, patch_hw[1] // 16)
        self.style_vae_encoder_module = VAEEncoderWithSkips(in_ch_for_style_encoder, z_ch_for_style, patch_hw)
        self.image_vae_decoder_module = VAEDecoderWithSkips(
            z_ch_for_style, self.char_text_encoder_module.rnn_output_dim, out_ch_for_image, patch_hw)

    def forward(self, image, mask, texts):
        x = torch.cat([image, mask], 1)
        if self.repaired:
            mu, logvar, _, pooled = self.style_vae_encoder_module(x, return_pooled=True)
        else:
            mu, logvar, skips = self.style_vae_encoder_module(x)
        z = reparameterize(mu, logvar)
        t = self.char_text_encoder_module(texts)
        dec = self.image_vae_decoder_module
        img = dec.forward_repaired(z, t, pooled) if self.repaired else dec(z, t, skips)
        return img, mu, logvar



class VGGPerceptual(nn.Module):
    """vae-gan.py:300-311 -- get_vgg_feat + perceptual_loss: L1 between VGG16 features[:16] of the ImageNet-normalised
    images.  ``features`` restates torchvision's ``vgg16().features[:16]`` (cfg 64,64,M,128,128,M,256,256,256; same
    Sequential indices, so a torchvision state_dict loads), checked against torchvision itself in
    tests/test_oracle_golden.py.  PARITY UNPINNED for the pretrained IMAGENET1K_V1 weights the reference downloads
    (not available offline): tests use seeded random weights."""

    def __init__(self):
        super().__init__()
        layers, cin = [], 3
        for v in (64, 64, "M", 128, 128, "M", 256, 256, 256):
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers).eval()
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1), persistent=False)

    def forward(self, fake, real):
        return F.l1_loss(self.features((fake - self.mean) / self.std), self.features((real - self.mean) / self.std))


# ------------------------------------------------------------------------------------------------
# vae-gan-oldv.py family: 3-level U-Net, gated skips, 4-row text map
# ------------------------------------------------------------------------------------------------
class CharacterTokenEncoderOldV(nn.Module):
    """vae-gan-oldv.py:74-148 -- Embedding -> biGRU -> Conv1d(k=3) -> adaptive avg pool to W/16 -> the row repeated
    ``target_feature_height`` (4) times -> + learned positional encoding (1, C, 4, W/16)."""

    def __init__(self, alphabet_str, emb_dim, rnn_hidden_dim, rnn_layers, target_feature_width, target_feature_height=4):
        super().__init__()
        self.lut = {ch: i + 1 for i, ch in enumerate(alphabet_str)}
        self.embedding = nn.Embedding(len(alphabet_str) + 1, emb_dim, padding_idx=0)
        self.rnn = nn.GRU(emb_dim, rnn_hidden_dim, num_layers=rnn_layers, batch_first=True, bidirectional=True,
                          dropout=0.1 if rnn_layers > 1 else 0)
        self.rnn_output_dim = 2 * rnn_hidden_dim
        self.target_feature_width, self.target_feature_height = target_feature_width, target_feature_height
        self.conv1d = nn.Conv1d(self.rnn_output_dim, self.rnn_output_dim, kernel_size=3, padding=1)
        self.register_parameter("pos_enc", nn.Parameter(
            torch.randn(1, self.rnn_output_dim, target_feature_height, target_feature_width) * 0.02))

    def tokenize(self, texts, max_len=60):
        idx = torch.zeros(len(texts), max_len, dtype=torch.long)
        for r, t in enumerate(texts):
            ids = [self.lut.get(ch, 0) for ch in t][:max_len]
            if ids:
                idx[r, :len(ids)] = torch.tensor(ids, dtype=torch.long)
        return idx

    def forward(self, texts, max_len=60):
        idx = self.tokenize(texts, max_len).to(self.embedding.weight.device)
        out, _ = self.rnn(self.embedding(idx))
        x = F.adaptive_avg_pool1d(self.conv1d(out.permute(0, 2, 1)), self.target_feature_width)
        x = x.unsqueeze(2).expand(-1, -1, self.target_feature_height, -1)
        return x + self.pos_enc                                            # (B, 2*hid, 4, W/16)


class VAEEncoderWithSkips3(nn.Module):
    """vae-gan-oldv.py:187-224 -- three double-conv levels (32/64/128) + bottleneck 256, heads with kernel (H/8, W/8)."""

    def __init__(self, in_ch, z_ch, patch_hw, skip_chans=(32, 64, 128), bottleneck_ch=256):
        super().__init__()
        self.e_conv1 = double_conv(in_ch, skip_chans[0])
        self.pool1 = nn.MaxPool2d(2, 2)
        self.e_conv2 = double_conv(skip_chans[0], skip_chans[1])
        self.pool2 = nn.MaxPool2d(2, 2)
        self.e_conv3 = double_conv(skip_chans[1], skip_chans[2])
        self.pool3 = nn.MaxPool2d(2, 2)
        self.bottleneck_conv = double_conv(skip_chans[2], bottleneck_ch)
        k = (patch_hw[0] // 8, patch_hw[1] // 8)
        self.mu_head = nn.Conv2d(bottleneck_ch, z_ch, kernel_size=k)
        self.logvar_head = nn.Conv2d(bottleneck_ch, z_ch, kernel_size=k)

    def forward(self, x):
        s1 = self.e_conv1(x)
        s2 = self.e_conv2(self.pool1(s1))
        s3 = self.e_conv3(self.pool2(s2))
        b = self.bottleneck_conv(self.pool3(s3))
        return self.mu_head(b), self.logvar_head(b), [s1, s2, s3]


class GatedSkipConnection(nn.Module):
    """vae-gan-oldv.py:226-231 -- skip * sigmoid(alpha), alpha (1, C, 1, 1) initialised to 0.3."""

    def __init__(self, channels, alpha_init=0.3):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1, channels, 1, 1) * alpha_init)

    def forward(self, skip_feat):
        return skip_feat * torch.sigmoid(self.alpha)


class VAEDecoderWithSpatialFiLM3(nn.Module):
    """vae-gan-oldv.py:235-320 -- ConvT(k=(H/8,1)) -> BN -> ReLU; 3 x [ConvT2x2 s2 -> cat(gated skip) -> FiLM ->
    double conv]; Conv1x1; Sigmoid.  The text map has 4 rows, so every FiLM interpolation is a true 2-D bilinear
    resize, and the bottleneck input is the map resized to (1, W/8)."""

    def __init__(self, z_ch, text_channels_in, out_ch_image, patch_h, patch_w, skip_chans=(32, 64, 128), bottleneck_ch=256):
        super().__init__()
        self.initial_h, self.initial_w = patch_h // 8, patch_w // 8
        self.skip_gates = nn.ModuleList([GatedSkipConnection(skip_chans[2]), GatedSkipConnection(skip_chans[1]),
                                         GatedSkipConnection(skip_chans[0])])
        self.bottleneck_proc = nn.Sequential(
            nn.ConvTranspose2d(z_ch + text_channels_in, bottleneck_ch, kernel_size=(self.initial_h, 1), stride=1, padding=0),
            *_bn_relu(bottleneck_ch))
        c = bottleneck_ch
        for i, s in enumerate((skip_chans[2], skip_chans[1], skip_chans[0]), start=1):
            setattr(self, f"up_tconv{i}", nn.ConvTranspose2d(c, s, kernel_size=2, stride=2))
            setattr(self, f"spatial_film{i}", SpatialFiLMLayer(text_channels_in, 2 * s))
            setattr(self, f"conv_block{i}", double_conv(2 * s, s))
            c = s
        self.final_image_conv = nn.Conv2d(skip_chans[0], out_ch_image, kernel_size=1)
        self.output_activation_fn = nn.Sigmoid()

    def forward(self, z, text_base, skips):
        t0 = F.interpolate(text_base, size=(1, self.initial_w), mode="bilinear", align_corners=False)
        x = self.bottleneck_proc(torch.cat([z.expand(-1, -1, 1, self.initial_w), t0], 1))
        for i in (1, 2, 3):
            x = getattr(self, f"up_tconv{i}")(x)
            x = torch.cat([x, self.skip_gates[i - 1](skips[3 - i])], 1)
            x = getattr(self, f"spatial_film{i}")(x, text_base)
            x = getattr(self, f"conv_block{i}")(x)
        return self.output_activation_fn(self.final_image_conv(x))


class VAEGAN_UNet_SpatialFiLM_OldV(nn.Module):
    """vae-gan-oldv.py:323-368 (the script calls it VAEGAN_UNet_SpatialFiLM too)."""

    def __init__(self, in_ch_style=4, z_ch_style=128, out_ch_img=3, alphabet_str_text=ALPHABET_STR,
                 char_emb_dim_text=128, char_rnn_hidden_dim_text=256, char_rnn_layers_text=2, patch_hw=(64, 448)):
        super().__init__()
        self.char_text_encoder_module = CharacterTokenEncoderOldV(
            alphabet_str_text, char_emb_dim_text, char_rnn_hidden_dim_text, char_rnn_layers_text, patch_hw[1] // 16, 4)
        self.style_vae_encoder_module = VAEEncoderWithSkips3(in_ch_style, z_ch_style, patch_hw)
        self.image_vae_decoder_module = VAEDecoderWithSpatialFiLM3(
            z_ch_style, self.char_text_encoder_module.rnn_output_dim, out_ch_img, patch_hw[0], patch_hw[1])

    def forward(self, image, mask, texts):
        mu, logvar, skips = self.style_vae_encoder_module(torch.cat([image, mask], 1))
        z = reparameterize(mu, logvar)
        text_base = self.char_text_encoder_module(texts)
        return self.image_vae_decoder_module(z, text_base, skips), mu, logvar
