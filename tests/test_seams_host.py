"""CPU: the HOST logic above the C ABI — module mirrors, the three drop-in seams, padding masks, position ids, chunked prefill,
KV-cache bookkeeping, forced routing — executed without a GPU by swapping `aria_b200.ops` for the oracle-backed stand-ins of
tests/standin_ops.py.  The scenarios ARE the GPU tests (imported and re-run with DEV = "cpu"), so the GPU suite exercises exactly
this host code on the real kernels."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import standin_ops  # noqa: E402
import test_gpu_dropin as D  # noqa: E402
import test_gpu_parity_full as P  # noqa: E402


@pytest.fixture()
def cpu_ops(monkeypatch):
    standin_ops.patch(monkeypatch)
    monkeypatch.setattr(P, "DEV", "cpu")
    monkeypatch.setattr(D, "DEV", "cpu")
    torch.set_grad_enabled(False)
    yield
    torch.set_grad_enabled(True)


def test_forced_routing_whole_model(cpu_ops):
    P.test_whole_model_with_oracle_routing_every_token_within_tolerance()


def test_padded_batches(cpu_ops):
    P.test_mirror_padded_batch_equals_unpadded_runs()


def test_chunked_prefill_and_overflow(cpu_ops):
    P.test_mirror_chunked_prefill_then_decode_and_cache_overflow()


def test_labels_and_rejected_arguments(cpu_ops):
    P.test_mirror_labels_loss_and_rejected_arguments()


@pytest.mark.parametrize("d,E,k,I,T", [(256, 8, 2, 512, 32), (256, 64, 6, 128, 300)])
def test_install_on_reference_moe_layer(cpu_ops, d, E, k, I, T):
    D.test_install_on_reference_moe_layer(d, E, k, I, T)


def test_reference_model_with_all_seams(cpu_ops):
    D.test_reference_model_with_all_seams_forward_and_cached_decode()


def test_hf_aria_with_seams(cpu_ops):
    D.test_hf_aria_forward_and_generate_with_seams()


def test_hf_aria_padded_batch(cpu_ops):
    D.test_hf_aria_padded_batch_through_generate()
