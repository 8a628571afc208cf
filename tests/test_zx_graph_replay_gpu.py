"""CUDA-graph replay of frames (engine.cu: frame_impl, D4_GRAPH=1) on the GPU.

Strict since round 2: besides bit-equal rollouts and equal launch counts the test asserts that frames really were REPLAYED
(d4_graph_replays), which a silent fall-back to direct launches would not satisfy."""
import pytest
import torch

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ CUDA-graph replay of frames
# D4_GRAPH=1 (read when the engine context is created): a frame is run directly the first time its key (B, t, ...) is seen,
# captured into a CUDA graph the second time, and replayed from then on (engine.cu: frame_impl).  Same kernels, same order,
# same staging arithmetic - the third rollout must reproduce the first one exactly, with the same launch count reported.

def test_cuda_graph_replay_reproduces_direct_frames(monkeypatch):
    import test_gpu_parity as G
    from dreamer4_b200 import _lib as L
    monkeypatch.setenv('D4_GRAPH', '1')
    model, sd = G._mid_model('tf32x3')
    lib = L.load()
    T, B = 5, 6                                          # 90 token rows: no atomics anywhere, the arithmetic is deterministic
    noise = G.to_cuda(G.make_noise(model.cfg, T, B, seed=5))
    flags = dict(return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
    runs, launches = [], []
    side = torch.cuda.Stream()                           # stream capture needs a non-default stream
    torch.cuda.synchronize()
    for _ in range(3):                                   # direct, capture + replay, replay
        l0 = lib.d4_launch_count()
        with torch.cuda.stream(side):
            runs.append(model.generate(T, batch_size=B, noise=noise, **flags))
        torch.cuda.synchronize()
        launches.append(lib.d4_launch_count() - l0)
    assert launches[0] == launches[1] == launches[2] > 0
    # rollout 1 ran directly, rollout 2 captured each frame and replayed it, rollout 3 replayed: 2 * T graph launches
    assert lib.d4_graph_replays(model._ctx) == 2 * T, f'{lib.d4_graph_replays(model._ctx)} graph replays, expected {2 * T}'
    for later in runs[1:]:
        assert torch.equal(later.actions.discrete, runs[0].actions.discrete)
        for name in ('latents', 'rewards', 'values', 'agent_embed'):
            assert torch.equal(getattr(later, name), getattr(runs[0], name)), name
        assert torch.equal(later.log_probs.discrete, runs[0].log_probs.discrete)
        assert torch.equal(later.old_action_unembeds.discrete, runs[0].old_action_unembeds.discrete)
