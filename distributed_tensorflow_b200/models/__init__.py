"""Model families: MNIST MLP (the reference's model), linear regression, tower MLP, ResNet-18."""
from .mnist_mlp import build_mnist_mlp, mnist_mlp_param_shapes  # noqa: F401
from .resnet import (build_resnet18_graph, resnet18_forward, resnet18_init, resnet18_loss,  # noqa: F401
                     resnet18_param_shapes)
