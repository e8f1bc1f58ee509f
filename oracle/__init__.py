"""TEST INFRASTRUCTURE: CPU oracle for the HeteroGIN hot path (see hgin_oracle.py header).
Never imported by gnn_link_prediction_b200/."""
