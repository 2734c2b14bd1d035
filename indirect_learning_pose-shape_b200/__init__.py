"""B200-native SMPL decode -> project -> mask -> part-seg / silhouette (keras_smpl hot path)."""
