from aqml_b200.distance import l2_distance, l1_distance, manhattan_distance  # noqa: F401
