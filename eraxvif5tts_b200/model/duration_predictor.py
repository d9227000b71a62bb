"""DurationPredictor with the reference's constructor, parameter names (state_dict keys) and forward signatures
(model/duration_predictor.py:4-66).  The modules only hold the parameters; the eval-mode forward is two kernels in
libf5b200.so (csrc/align.cu: f5b_duration_predictor).  Training it (p_dropout, backward) is not built."""
import torch
import torch.nn as nn

from .. import _lib as L

f32 = torch.float32


class DurationPredictor(nn.Module):
    def __init__(self, text_num_embeds, in_channels, filter_channels, kernel_size, p_dropout, gin_channels=0):
        super().__init__()
        if kernel_size % 2 != 1:
            raise NotImplementedError("odd kernel sizes only (padding = kernel_size // 2 keeps the length)")
        self.text_embed = nn.Embedding(text_num_embeds + 1, in_channels)  # 0 is the filler token
        self.in_channels, self.filter_channels, self.kernel_size = in_channels, filter_channels, kernel_size
        self.p_dropout, self.gin_channels = p_dropout, gin_channels
        self.drop = nn.Dropout(p_dropout)
        self.conv_1 = nn.Conv1d(in_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_1 = nn.GroupNorm(1, filter_channels)
        self.conv_2 = nn.Conv1d(filter_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_2 = nn.GroupNorm(1, filter_channels)
        self.proj = nn.Conv1d(filter_channels, 1, 1)
        if gin_channels != 0:
            self.cond = nn.Conv1d(gin_channels, in_channels, 1)

    def _run(self, ids, mask, id_shift, g):
        if g is not None:
            raise NotImplementedError("speaker conditioning g (gin_channels) is unused by the reference's call sites and is not built")
        dev = self.proj.weight.device
        if dev.type != "cuda":
            raise L.F5bError("DurationPredictor needs its parameters on a CUDA device (B200); there is no CPU fallback")
        if self.training and self.p_dropout > 0:
            raise NotImplementedError("only the eval-mode forward is built: call .eval()")
        ids = ids.to(device=dev, dtype=torch.int64).contiguous()
        mask = mask.to(device=dev, dtype=f32).contiguous()
        b, nt = ids.shape
        lo, hi = int(ids.min()) + id_shift, int(ids.max()) + id_shift
        if lo < 0 or hi >= self.text_embed.weight.shape[0]:
            raise IndexError("index out of range in self")  # nn.Embedding's error
        p = [t.detach().to(f32).contiguous() for t in (self.text_embed.weight, self.conv_1.weight, self.conv_1.bias, self.norm_1.weight,
                                                       self.norm_1.bias, self.conv_2.weight, self.conv_2.bias, self.norm_2.weight,
                                                       self.norm_2.bias, self.proj.weight, self.proj.bias)]
        h1 = torch.empty(b, nt, self.filter_channels, dtype=f32, device=dev)
        out = torch.empty(b, 1, nt, dtype=f32, device=dev)
        L.check(L.load().f5b_duration_predictor(ids.data_ptr(), id_shift, mask.data_ptr(), p[0].data_ptr(), p[0].shape[0],
                                                *[t.data_ptr() for t in p[1:]], h1.data_ptr(), out.data_ptr(), b, nt, self.in_channels,
                                                self.filter_channels, self.kernel_size, L.stream()), "f5b_duration_predictor")
        return out

    @torch.no_grad()
    def forward(self, x, x_mask, g=None):
        """x int [b, nt] text tokens padded with -1 (list_str_to_idx), x_mask [b, nt] -> log-durations [b, 1, nt]"""
        return self._run(x, x_mask, 1, g)

    @torch.no_grad()
    def phoneme_forward(self, phoneme_indices, phoneme_mask, g=None):
        """same network on phoneme indices (no +1 shift), duration_predictor.py:46-66"""
        return self._run(phoneme_indices, phoneme_mask, 0, g)
