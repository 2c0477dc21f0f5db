"""Import-path alias: ``qpwcnet.core.{layers,non_layers,warp,util}`` resolve to the B200-native
implementation in ``qpwcnet_b200.core`` so that code written against the reference's module paths
(test/test_cost_volume.py:4, test/test_warp.py:6-7) runs unchanged."""
