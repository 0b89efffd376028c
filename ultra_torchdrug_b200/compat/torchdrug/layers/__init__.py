"""`torchdrug.layers` stand-in: MessagePassingBase + MLP (reference layer.py:14,193,228; model.py:53)."""
from torch import nn
from torch.nn import functional as F

from . import functional  # noqa: F401


class MessagePassingBase(nn.Module):
    """forward = combine(input, message_and_aggregate(graph, input)); the default
    message_and_aggregate is aggregate(message()) - the fallback the reference reaches through
    `super()` at layer.py:113,300."""

    gradient_checkpoint = False

    def message(self, graph, input):
        raise NotImplementedError

    def aggregate(self, graph, message):
        raise NotImplementedError

    def message_and_aggregate(self, graph, input):
        message = self.message(graph, input)
        update = self.aggregate(graph, message)
        return update

    def combine(self, input, update):
        raise NotImplementedError

    def forward(self, graph, input):
        update = self.message_and_aggregate(graph, input)
        output = self.combine(input, update)
        return output


class MLP(nn.Module):
    """Linear stack with activation between layers; parameters live under `layers.{i}.*`."""

    def __init__(self, input_dim, hidden_dims, short_cut=False, batch_norm=False, activation="relu", dropout=0):
        super(MLP, self).__init__()
        if not isinstance(hidden_dims, (list, tuple)):
            hidden_dims = [hidden_dims]
        self.dims = [input_dim] + list(hidden_dims)
        self.short_cut = short_cut
        self.activation = getattr(F, activation) if isinstance(activation, str) else activation
        self.dropout = nn.Dropout(dropout) if dropout else None
        self.layers = nn.ModuleList(nn.Linear(self.dims[i], self.dims[i + 1]) for i in range(len(self.dims) - 1))
        self.batch_norms = nn.ModuleList(nn.BatchNorm1d(d) for d in self.dims[1:-1]) if batch_norm else None

    def forward(self, input):
        layer_input = input
        for i, layer in enumerate(self.layers):
            hidden = layer(layer_input)
            if i < len(self.layers) - 1:
                if self.batch_norms:
                    hidden = self.batch_norms[i](hidden.flatten(0, -2)).view_as(hidden)
                hidden = self.activation(hidden)
                if self.dropout:
                    hidden = self.dropout(hidden)
            if self.short_cut and hidden.shape == layer_input.shape:
                hidden = hidden + layer_input
            layer_input = hidden
        return hidden
