"""-m gpu parity tests: CUDA path (through the C ABI) vs the CPU oracle and the golden fixtures."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _cases():
    import gpu_cases
    return gpu_cases


def _check(metrics):
    bad = {k: v for k, v in metrics.items() if not (v[0] <= v[1])}
    assert not bad, f"out of tolerance: {bad} (all: {metrics})"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(cuda_device):
    from protein_redesign_b200 import _lib
    assert _lib.load().prd_device_check() == 0, _lib.last_error()


CASE_NAMES = [
    "gemm_basic", "gemm_k512", "gemm_tails", "gemm_n48", "gemm_batch", "gemm_epilogue", "gemm_fp16_out", "gemm_fp16_relu", "gemm_fp16_sigmoid",
    "pair_transition", "pair_transition_readme", "pair_transition_n140", "trimul_outgoing", "trimul_incoming",
    "trimul_readme", "trimul_n140", "trimul_n300", "trimul_n256", "trimul_n128", "triattn_starting", "triattn_ending", "triattn_n200", "triattn_n140",
    "triattn_n300", "triattn_n512", "triattn_readme", "outer_linear", "outer_linear_readme", "outer_linear_n300",
    "single_attention", "single_transition", "spattention", "opm", "embeddings", "embeddings_readme", "embeddings_n128", "embeddings_n256", "heads",
    "pair_bias", "sample_eager", "sample_graph", "invariants", "invariants_n300", "invariants_n512_b8", "invariants_n1024", "sample_graph_T50", "loss_paper_n72",
    # round 2: oracle parity at BASELINE.json's sizes, public module entry points, advisor regressions
    "step_n512", "step_n512_b2_ragged", "step_n300", "step_n1024", "batch_rows_b8_n512", "trimul_n512_outgoing", "trimul_n512_incoming",
    "triattn_n512_ending", "outer_linear_n512", "pair_transition_n512", "heads_n512", "embeddings_n512", "denoiser_forward",
    "folding_block_forward", "predict_step_ema", "back_to_back_batches",
    # round 2: backward pass (every prd_<op>_bwd vs autograd through the oracle)
    "gemm_tf32", "gemm_tf32_batch_tails", "gemm_tf32_epilogue", "gemm_tf32_epilogue_unaligned", "attn_tc_vs_simt", "attn_tc_vs_simt_starting", "bwd_pair_fc", "bwd_single_fc", "bwd_triattn_starting", "bwd_triattn_ending", "bwd_triattn_n140",
    "bwd_trimul_outgoing", "bwd_trimul_incoming", "bwd_trimul_n75", "bwd_outer_linear", "bwd_single_attention", "bwd_spattention",
    "bwd_heads", "bwd_embeddings", "bwd_embeddings_readme", "train_step_paper_n72",
    # round 2: Lightning-free predict loop and GPU post-processing
    "postprocess", "predict_loop", "train_modes", "dw_tc", "mask_contract", "fused_bias", "fused_bias_readme",
]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_case(name):
    _check(_cases().CASES[name]())


@pytest.mark.parametrize("tag,cfg_name,sizes,seed", [
    ("paper_n72", "PAPER", ((12, 60), (9, 50)), 3),
    ("readme_n40", "README", ((8, 32), (6, 27)), 2),
    ("paper_n128", "PAPER", ((16, 112),), 4),
])
def test_step_vs_oracle_and_reference_golden(tag, cfg_name, sizes, seed):
    gc = _cases()
    from protein_redesign_b200 import synthetic as syn
    gold = load_golden(f"step_{tag}.npz")
    _check(gc.case_step(getattr(syn, cfg_name), sizes, seed=seed, golden=gold, probes=(tag == "paper_n72")))


@pytest.mark.parametrize("tag,cfg_name,over,sizes,seed,kw", [
    ("readme_n40", "README", dict(mask_prob=0.15, num_steps=2000), ((8, 32), (6, 27)), 10, {}),
])
def test_training_objective_vs_oracle_and_reference_golden(tag, cfg_name, over, sizes, seed, kw):
    import dataclasses
    gc = _cases()
    from protein_redesign_b200 import synthetic as syn
    cfg = dataclasses.replace(getattr(syn, cfg_name), **over)
    _check(gc.case_loss(cfg, sizes, seed, golden=load_golden(f"loss_{tag}.npz"), **kw))


def test_training_step_gradients_vs_oracle_and_reference_fingerprints():
    """loss.backward() through the CUDA backward kernels: all 240 parameter gradients against autograd through the oracle
    and against the unmodified reference's gradient fingerprints (tests/golden/loss_readme_n40.npz)."""
    import dataclasses
    gc = _cases()
    from protein_redesign_b200 import synthetic as syn
    cfg = dataclasses.replace(syn.README, mask_prob=0.15, num_steps=2000)
    _check(gc.case_train_step(cfg, ((8, 32), (6, 27)), 10, golden=load_golden("loss_readme_n40.npz")))


def test_cpu_tensor_is_rejected():
    """No CPU fallback: the product path must fail loudly on a CPU tensor."""
    from protein_redesign_b200 import ops
    from protein_redesign_b200 import synthetic as syn
    x = torch.zeros(1, 8, 8, 64)
    with pytest.raises(RuntimeError):
        ops.symmetrize(syn.PAPER, x)
