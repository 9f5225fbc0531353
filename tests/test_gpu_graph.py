"""CUDA-graph replay of the generator's forward + backward (sparenet_b200/graph.py) must reproduce the eager step: every
sm_100a kernel on the path has to be capture-safe (current stream only, scratch from the caching allocator, no syncs)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_forward_backward_matches_eager(cuda):
    from sparenet_b200.dropin.utils.model_init import init_weights
    from sparenet_b200.dropin.cuda.chamfer_distance import ChamferDistanceMean
    from sparenet_b200.dropin.models.sparenet_generator import SpareNetGenerator
    from sparenet_b200.graph import GraphedForwardBackward
    # fp32 library GEMMs for this comparison: under capture cuBLAS/cuDNN may pick other TF32 kernels (split-K, workspace), whose
    # 1e-3 rounding differences would otherwise dominate what is compared here
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    net = SpareNetGenerator(n_primitives=8, hide_size=256, bottleneck_size=256, num_points=8 * 512, use_SElayer=True, use_AdaIn="share",
                            encode="Residualnet")
    net.apply(init_weights)
    net = net.to(cuda).train()
    state = copy.deepcopy(net.state_dict())
    cd = ChamferDistanceMean()
    partial = torch.rand(4, 1024, 3, device=cuda) - 0.5
    gt = torch.rand(4, 4096, 3, device=cuda) - 0.5

    def loss_fn(p, g):
        coarse, middle, refine, loss_mst = net({"partial_cloud": p})
        return cd(coarse, g).mean() + cd(middle, g).mean() + cd(refine, g).mean() + 0.1 * loss_mst.mean()

    params = [p for p in net.parameters()]
    loss_e = loss_fn(partial, gt)
    loss_e.backward()
    grads_e = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    loss_e = loss_e.item()
    del cd                                                   # nothing may keep the eager graph alive across the capture
    cd = ChamferDistanceMean()
    for p in params:
        p.grad = None
    net.load_state_dict(state)                               # same BatchNorm running statistics as the eager step saw
    g = GraphedForwardBackward(loss_fn, params, (partial, gt), warmup=2)
    net.load_state_dict(state)
    loss_g = g(partial, gt).item()
    assert abs(loss_g - loss_e) <= 1e-6 * abs(loss_e)
    # Parameters whose gradient is identically zero in exact arithmetic (conv biases in front of a normalisation) carry
    # only rounding noise, which float atomics reorder between runs: each deviation is measured against the parameter's
    # own gradient scale plus a floor of 1e-4 of the largest gradient in the model.
    gmax = max(v.abs().max().item() for v in grads_e.values())
    rows = []
    for n, p in net.named_parameters():
        if n in grads_e:
            scale = grads_e[n].abs().max().item() + 1e-4 * gmax
            rows.append(((p.grad - grads_e[n]).abs().max().item() / scale, n, grads_e[n].abs().max().item()))
    rows.sort(reverse=True)
    worst = rows[0][0]
    print(f"[graph] loss eager {loss_e:.8f} graph {loss_g:.8f}; largest gradient {gmax:.3e}; worst deviations / scale:")
    for r in rows[:5]:
        print(f"    {r[0]:.2e}  {r[1]}  (|grad|max {r[2]:.3e})")
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    # float atomics in the scatter kernels reorder additions between runs, and a re-rounded near-tie can hand a max/min pooling
    # winner to its neighbour: deviations stay at the 1e-3 level of the gradient scale for the deepest (encoder) parameters
    assert worst < 5e-3, rows[:5]
    loss_g2 = g(partial, gt).item()                          # replay is repeatable
    assert abs(loss_g2 - loss_g) <= 1e-6 * abs(loss_g)


@pytest.mark.gpu
def test_grad_arena_pack_one_launch_eager_and_captured(cuda):
    """GradArena.pack (own_grads=False) through snb_multi_copy: odd sizes, a 16-byte-misaligned gradient view, several eager calls
    (alternating pinned pointer tables) and one call captured in a CUDA graph and replayed after the gradients changed in place."""
    from sparenet_b200.dist import GradArena
    torch.manual_seed(5)
    shapes = [(5, 3), (7,), (300, 41), (1,), (4096,), (4097,), (2, 2, 2)]
    ps = [torch.nn.Parameter(torch.randn(*s, device=cuda)) for s in shapes]
    arena = GradArena(ps, own_grads=False)
    big = torch.randn(10000, device=cuda)
    for round_ in range(3):
        for i, p in enumerate(ps):
            p.grad = None if i == 3 else torch.full_like(p, float(10 * round_ + i + 1))
        ps[1].grad = big[1 + round_:8 + round_]               # a view whose address is not a multiple of 16
        arena.pack()
        torch.cuda.synchronize()
        for i, p in enumerate(ps):
            want = torch.zeros_like(p) if i == 3 else (big[1 + round_:8 + round_] if i == 1 else torch.full_like(p, float(10 * round_ + i + 1)))
            assert torch.equal(arena.views[id(p)], want), (round_, i)
    # captured: static gradient tensors, values change between replays
    static = [torch.zeros_like(p) for p in ps]
    for p, g in zip(ps, static):
        p.grad = g
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        arena.pack()
    for val in (3.0, -7.5):
        for g in static:
            g.fill_(val)
        arena.pack()                                          # an eager call in between must not disturb the captured table
        graph.replay()
        torch.cuda.synchronize()
        for p in ps:
            assert torch.equal(arena.views[id(p)], torch.full_like(p, val))
