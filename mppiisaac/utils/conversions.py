from mppi_isaac_b200.utils.conversions import matrix_to_euler_angles, quaternion_to_matrix, quaternion_to_yaw  # noqa: F401
