"""Import-path glue so that the reference's own entry points find the B200 implementation
(`from src.geometry import compute_streamline_metrics`, /root/reference/README.md:78)."""
