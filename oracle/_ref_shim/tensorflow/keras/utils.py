_CUSTOM = {}


def get_custom_objects():
    return _CUSTOM
