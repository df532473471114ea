"""UNET1D fixture configurations shared by make_golden_unet.py and tests/test_oracle_golden.py."""
from after_b200 import config

CASES = {
    # per-scale time_cond embedding, attention in the deepest encoder block, the middle block and the first decoder block
    "attn": (config.UNetConfig(in_size=16, channels=[32, 64, 64], ratios=[2, 2], kernel_size=5, time_channels=32,
                               time_cond_in_channels=4, time_cond_channels=16, cond_channels=6, n_attn_layers=1), 91),
    # time_cond concatenated to the input, a ratio-1 stage, residual on the last block, no global condition
    "concat": (config.UNetConfig(in_size=8, out_size=8, channels=[16, 32, 32], ratios=[1, 2], kernel_size=3, time_channels=16,
                                 time_cond_in_channels=4, time_cond_channels=0, cond_channels=0, n_attn_layers=0,
                                 use_res_last=True), 92),
}
