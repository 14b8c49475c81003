"""Fused group-product scan (cumprod / cummul on group LieTensors) — filled in by csrc/scan.cu."""


def try_cumprod(input, dim, left):
    return None
