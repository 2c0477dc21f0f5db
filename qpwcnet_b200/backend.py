"""Process-global image data format, standing in for ``tf.keras.backend.image_data_format()``
which the reference layers read at construction (qpwcnet/core/layers.py:41,119,146,173)."""
_FORMAT = "channels_last"          # Keras' default; the reference's training scripts set channels_first


def image_data_format() -> str:
    return _FORMAT


def set_image_data_format(data_format: str) -> None:
    global _FORMAT
    get_axis(data_format)
    _FORMAT = data_format


def get_axis(data_format: str) -> int:
    """Channel axis of a data format (reference: ``_get_axis``, qpwcnet/core/layers.py:19-29)."""
    if data_format == "channels_first":
        return 1
    if data_format == "channels_last":
        return 3
    raise ValueError("Unsupported data format : {}".format(data_format))
