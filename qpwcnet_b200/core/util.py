"""``disable_gpu`` shim (qpwcnet/core/util.py:13-27).  The reference's test/test_warp.py:11 calls it
to force TF onto the CPU; this package has no CPU path, so it is a documented no-op."""


def disable_gpu():
    return None
