"""Small non-Hugging-Face transformer-ish models used by the absorb-discovery and AWQ folding tests.  The reference's
torch.jit tracer still works on plain nn.Modules like these (it does not on transformers >= 5 models), so they are the
models on which the live reference can generate fixtures for absorption."""
import torch


class Block(torch.nn.Module):
    def __init__(self, d=32, variant=0):
        super().__init__()
        self.variant = variant
        self.ln1 = torch.nn.LayerNorm(d)
        self.q = torch.nn.Linear(d, d)
        self.k = torch.nn.Linear(d, d)
        self.v = torch.nn.Linear(d, d)
        self.o = torch.nn.Linear(d, d)
        self.ln2 = torch.nn.LayerNorm(d)
        self.fc1 = torch.nn.Linear(d, 2 * d)
        self.fc2 = torch.nn.Linear(2 * d, d)
        self.act = torch.nn.LeakyReLU(0.1) if variant == 1 else torch.nn.ReLU()

    def forward(self, x):
        h = self.ln1(x)
        a = torch.softmax(self.q(h) * self.k(h), dim=-1) * self.v(h)
        x = x + self.o(a)
        h = self.ln2(x)
        if self.variant == 1:      # the norm output also feeds a residual add: nothing may be folded into ln2
            return x + h + self.fc2(self.act(self.fc1(h)))
        if self.variant == 2:      # cast + element-wise mul consumers
            g = self.fc1(h.to(torch.float32))
            return x + self.fc2(self.act(g) * 0.5)
        if self.variant == 3:      # a view between producer and consumer
            return x + self.fc2(self.act(self.fc1(h)).reshape(x.shape[0], -1, 2 * x.shape[-1]))
        return x + self.fc2(self.act(self.fc1(h)))


class Toy(torch.nn.Module):
    def __init__(self, d=32, n=2, variant=0, vocab=64):
        super().__init__()
        self.emb = torch.nn.Embedding(vocab, d)
        self.layers = torch.nn.ModuleList([Block(d, variant) for _ in range(n)])
        self.norm = torch.nn.LayerNorm(d)
        self.head = torch.nn.Linear(d, vocab)

    def forward(self, ids):
        x = self.emb(ids)
        for b in self.layers:
            x = b(x)
        return self.head(self.norm(x))
