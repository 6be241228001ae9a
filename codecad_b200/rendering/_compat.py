"""Nothing to adapt yet: shapes only need dimension(), bounding_box() and feature_size()."""
