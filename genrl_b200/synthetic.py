"""Seeded synthetic replay batches with the reference's `data` dict contract (tools/replay.py:178-236):
observation u8 [B,T,3,64,64], action f32 [B,T,A] in (-1,1), reward f32 [B,T,1], discount ones, is_first[:,0]=True (+ optional random
resets), is_last/is_terminal False, clip_video unit-norm f32 [B,T,clip_dim].  Draw order is fixed so the same seed gives the same batch
as the test-side generator (tests assert that)."""
import torch
import torch.nn.functional as F


def synthetic_batch(B, T, A, clip_dim=512, seed=0, resets=0.0, img=64):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randint(0, 256, (B, T, 3, img, img), generator=g, dtype=torch.uint8)
    action = torch.rand(B, T, A, generator=g) * 2 - 1
    reward = torch.rand(B, T, 1, generator=g) * 2
    is_first = torch.zeros(B, T, dtype=torch.bool)
    is_first[:, 0] = True
    if resets > 0:
        is_first |= torch.rand(B, T, generator=g) < resets
    clip = F.normalize(torch.randn(B, T, clip_dim, generator=g), dim=-1)
    return dict(observation=obs, action=action, reward=reward, discount=torch.ones(B, T, 1), is_first=is_first,
                is_last=torch.zeros(B, T, dtype=torch.bool), is_terminal=torch.zeros(B, T, dtype=torch.bool), clip_video=clip)
