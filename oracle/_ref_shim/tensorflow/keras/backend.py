_FORMAT = ["channels_last"]            # Keras' default


def image_data_format():
    return _FORMAT[0]


def set_image_data_format(fmt):
    if fmt not in ("channels_first", "channels_last"):
        raise ValueError("Unknown data_format: " + str(fmt))
    _FORMAT[0] = fmt
