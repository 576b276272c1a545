"""Test infrastructure: the oracle (CPU restatement of the reference hot path). Never imported by the product."""
