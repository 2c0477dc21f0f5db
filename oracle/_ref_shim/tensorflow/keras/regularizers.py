def l2(l2=0.01):  # noqa: A002
    return ("l2", l2)
