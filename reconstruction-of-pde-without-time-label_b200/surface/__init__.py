"""Drop-in module surface: the reference's nn.Module classes, same constructor signatures,
parameter names / shapes / dtypes / registration order, backed by the CUDA ops."""
