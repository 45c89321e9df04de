"""`src.geometry` as the reference's README advertises it (/root/reference/README.md:78-81; the
reference's own src/geometry/__init__.py is empty, so that import fails there — SURVEY.md F3)."""
from lesion_condition_vae_b200.tract_geom_proc import (  # noqa: F401
    compute_streamline_metrics, compute_streamline_metrics_csr, read_streamlines_from_vtk,
)
