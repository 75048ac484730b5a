"""HOST logic of the CLI / scan API on a machine without a GPU: the device entry points are
replaced by the CPU oracle (tests/oracle_backend.py) so that packing, hit mapping, frame
assembly, messages and TSV text are compared with the reference's golden outputs.  The same
cases run through the real kernels in tests/test_cli_gpu.py (-m gpu)."""
import warnings

import pytest

import test_cli_gpu as gpu_cases
from oracle_backend import install

FUNCS = [getattr(gpu_cases, n) for n in dir(gpu_cases) if n.startswith("test_")
         and n not in ("test_cli_output_is_byte_identical", "test_pwm_module_signature_and_errors")]


@pytest.mark.parametrize("name", gpu_cases.ALIGNED)
def test_cli_host_logic_against_golden(name, in_repo, monkeypatch):
    install(monkeypatch)
    gpu_cases.test_cli_output_is_byte_identical.__wrapped__(name, in_repo) \
        if hasattr(gpu_cases.test_cli_output_is_byte_identical, "__wrapped__") else \
        _call(gpu_cases.test_cli_output_is_byte_identical, name=name, in_repo=in_repo)


def _call(fn, **available):
    import inspect
    names = inspect.signature(fn).parameters
    return fn(**{k: available[k] for k in names})


@pytest.mark.parametrize("fn", FUNCS, ids=[f.__name__ for f in FUNCS])
def test_api_host_logic_against_golden(fn, in_repo, golden_api, capsys, monkeypatch):
    install(monkeypatch)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _call(fn, in_repo=in_repo, golden_api=golden_api, capsys=capsys)
