"""Drop-in call surface of ``qpwcnet.core`` for the cost-volume / warp path (layers, functors,
``tf_warp``), backed by libqpwc."""
