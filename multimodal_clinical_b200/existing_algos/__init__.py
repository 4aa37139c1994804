"""Device-resident counterparts of the reference's existing_algos/ package (QMF, OGM-GE)."""
