from x_mlps_pytorch.mlp import MLP, create_mlp
