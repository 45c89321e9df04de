"""Drop-in module: same import path and public names as
/root/reference/src/geometry/tract_geom_proc.py, backed by the CUDA path.  The reference driver
imports it flat (`from tract_geom_proc import compute_streamline_metrics`,
comprehensive_tract_geometry_analysis.py:22): put this directory on sys.path, or copy this stub
next to the driver."""
from lesion_condition_vae_b200.tract_geom_proc import *  # noqa: F401,F403
from lesion_condition_vae_b200.tract_geom_proc import (  # noqa: F401
    compute_streamline_metrics, compute_streamline_metrics_csr, read_streamlines_from_vtk,
)
