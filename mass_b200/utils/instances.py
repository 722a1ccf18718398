"""Instance extraction for SemanticProjectionLayer.find (filled in with the K3 kernels)."""


def find_instances(layer, semantic_category, confidence_threshold, contour_padding, contour_threshold,
                   feature_map):
    raise NotImplementedError("find(): instance pooling kernels not built yet")
