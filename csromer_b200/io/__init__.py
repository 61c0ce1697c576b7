from .io_functions import DeviceCube, Reader, Writer, filter_cubes

__all__ = ["Reader", "Writer", "filter_cubes", "DeviceCube"]
