"""altformer_b200: B200-native hot path of ST-GCN-AltFormer.  Import it as `altformer_b200` (see the
shim package of that name at the repository root)."""
