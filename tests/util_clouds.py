"""Thin re-export of the seeded synthetic generators (unopose_b200/synthetic.py)."""
from unopose_b200.synthetic import (batch_clouds, matching_batch, matching_instance, random_rotation,  # noqa: F401
                                    surface_cloud, unit_cloud)
