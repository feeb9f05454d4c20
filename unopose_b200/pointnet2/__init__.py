"""Drop-in for the reference package core.unopose.model.pointnet2 (hot-path ops only)."""
