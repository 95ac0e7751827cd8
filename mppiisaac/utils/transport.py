from mppi_isaac_b200.utils.transport import bytes_to_torch, torch_to_bytes  # noqa: F401
