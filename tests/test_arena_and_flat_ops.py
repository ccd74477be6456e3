

def test_arena_carved_from_one_segment_keeps_batchnorm_backward_valid() -> None:
    """SPMD arenas are carved out of ONE symmetric-memory segment: every parameter, buffer and companion region is a
    view of the same base tensor.  Views share the base's autograd version counter, so one BatchNorm layer's in-place
    running-statistics update used to invalidate what another layer had saved for backward (the client's fit failed
    with "modified by an inplace operation" and -- with accept_failures -- the federation silently stopped learning).
    Buffers are re-homed behind tensors with their own counters."""
    import torch
    from torch import nn

    from fl4health_b200.parallel.arena import attach_arena

    segment, cursor = torch.zeros(1 << 20, dtype=torch.uint8), [0]

    def carve(numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        start, cursor[0] = cursor[0], cursor[0] + (nbytes + 255) // 256 * 256
        return segment[start:start + nbytes].view(dtype)

    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 8, 3), nn.BatchNorm2d(8), nn.ReLU(),
                        nn.Conv2d(8, 8, 3), nn.BatchNorm2d(8))
    arena = attach_arena(net, allocator=carve)
    momentum = arena.companion("momentum")  # a companion region from the same segment, updated in place between steps
    for _ in range(2):
        out = net(torch.randn(4, 3, 12, 12))
        momentum.add_(1.0)
        out.sum().backward()  # used to raise
    assert net[1].running_mean.data_ptr() == arena.view("1.running_mean").data_ptr()  # still the arena's memory
    assert float(arena.view("4.running_var").sub(1).abs().sum()) > 0  # and the updates landed there
    assert int(net[1].num_batches_tracked) == 2
