"""Fourier feature embedder (reference: networks/embedder.py:5-54): [x, sin(x 2^k), cos(x 2^k)]_{k<num_freqs}.
Weight-less; the arithmetic runs in dd_fourier_embed / dd_box_features (fp32)."""


class Embedder:
    def __init__(self, input_dims, num_freqs, include_input=True, log_sampling=True):
        if not include_input or not log_sampling:
            raise NotImplementedError("only include_input=True, log_sampling=True (the reference's configs)")
        self.input_dims, self.num_freqs = input_dims, num_freqs
        self.out_dim = input_dims * (1 + 2 * num_freqs)

    def __call__(self, inputs):
        from .. import ops
        assert inputs.shape[-1] == 3
        flat = inputs.reshape(-1, 3).float().contiguous()
        return ops.fourier_embed(flat, self.num_freqs).reshape(*inputs.shape[:-1], self.out_dim)


def get_embedder(input_dims, num_freqs, include_input=True, log_sampling=True):
    return Embedder(input_dims, num_freqs, include_input, log_sampling)
