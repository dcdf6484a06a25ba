"""The lines of AsyncTokenProcess::GetTimingInfoAndResetTimer (async_token_process.h:273-414) as regular expressions,
shared by the CPU test that reads them off the reference's own output (tests/test_oracle_background.py) and the GPU test
that holds the drop-in module's report to them (tests/test_python_api_gpu.py)."""
import re

NUM = r"\d+ ms \(\d+ (batches|tokens); \d+ ms avg\)"
LINES = (("Batch loading", "on time between each generated batch"), ("Batch gen", "on generating batches"),
         ("Result consume", "on handling results"), (r"Unit \[1\]", "on ingesting tokens in workers"))


def check(out: str):
    for head, tail in LINES:
        assert re.search(rf"^{head}: {NUM} {tail}$", out, re.M), (head, out)
