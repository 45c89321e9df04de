"""Drop-in module name of the reference driver
(/root/reference/src/geometry/comprehensive_tract_geometry_analysis.py), backed by the batched
B200 driver."""
from lesion_condition_vae_b200.tract_driver import (  # noqa: F401
    TRACT_LIST, generate_summary_statistics, get_all_subjects, load_config, main, process_all_tracts,
    process_single_tract,
)
