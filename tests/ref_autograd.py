"""torch.func restatement of the reference's kernel functionals, structure 1:1 with
``models/GP.py:28-180`` (nested autodiff, rotated-coordinate 5-index "Laplacian"),
in float64 and without the float16 casts.  Used only to pin the oracle's closed forms.
"""
import torch
from torch.func import grad


class RefKernels:
    def __init__(self, d, idx_set, sigma_eq=0.25):
        self.d = d
        self.sigma = sigma_eq * (d ** 0.5)                 # models/GP.py:25
        self.idx = [int(i) for i in idx_set]

    def laplacian_op(self, f):                             # models/GP.py:28-39
        def hvp(f, x, i):
            call_jei = lambda x: grad(f)(x)[i]
            return grad(call_jei)(x)

        def laplacian(x):
            vals = [hvp(f, x, i)[i] for i in self.idx]
            return torch.stack(vals).mean() * self.d
        return laplacian

    def kappa(self, x, y):                                 # :41-43
        return torch.exp(-torch.sum((x - y) ** 2) / (2 * self.sigma ** 2))

    def dx_t_kappa(self, x, y):
        return grad(self.kappa, argnums=0)(x, y)

    def dt_x_t_kappa(self, x, y):
        return self.dx_t_kappa(x, y)[-1]

    def dy_t_kappa(self, x, y):
        return grad(self.kappa, argnums=1)(x, y)

    def dt_y_t_kappa(self, x, y):
        return self.dy_t_kappa(x, y)[-1]

    def div_x_kappa(self, x, y):
        return torch.sum(self.dx_t_kappa(x, y)[:-1])

    def div_y_kappa(self, x, y):
        return torch.sum(self.dy_t_kappa(x, y)[:-1])

    def laplacian_x_t_kappa(self, x_t, y_t):               # :87-95
        t_x, x = x_t[0:1], x_t[1:]
        f = lambda x: self.kappa(torch.cat((x, t_x)), y_t)
        return self.laplacian_op(f)(x)

    def laplacian_y_t_kappa(self, x_t, y_t):               # :97-105
        t_y, y = y_t[0:1], y_t[1:]
        f = lambda y: self.kappa(x_t, torch.cat((y, t_y)))
        return self.laplacian_op(f)(y)

    def dt_x_t_dt_y_t_kappa(self, x, y):                   # :107-111
        return grad(self.dt_x_t_kappa, argnums=1)(x, y)[-1]

    def dt_x_t_div_y_kappa(self, x, y):                    # :113-117
        return torch.sum(grad(self.dt_x_t_kappa, argnums=1)(x, y)[:-1])

    def dt_x_t_laplacian_y_t_kappa(self, x_t, y_t):        # :119-127
        t_y, y = y_t[0:1], y_t[1:]
        f = lambda y: self.dt_x_t_kappa(x_t, torch.cat((y, t_y)))
        return self.laplacian_op(f)(y)

    def div_x_dt_y_t_kappa(self, x, y):                    # :129-133
        return grad(self.div_x_kappa, argnums=1)(x, y)[-1]

    def div_x_div_y_kappa(self, x, y):                     # :135-139
        return torch.sum(grad(self.div_x_kappa, argnums=1)(x, y)[:-1])

    def div_x_laplacian_y_t_kappa(self, x_t, y_t):         # :141-149
        t_y, y = y_t[0:1], y_t[1:]
        f = lambda y: self.div_x_kappa(x_t, torch.cat((y, t_y)))
        return self.laplacian_op(f)(y)

    def laplacian_x_t_dt_y_t_kappa(self, x_t, y_t):        # :151-159
        t_x, x = x_t[0:1], x_t[1:]
        f = lambda x: self.dt_y_t_kappa(torch.cat((x, t_x)), y_t)
        return self.laplacian_op(f)(x)

    def laplacian_x_t_div_y_kappa(self, x_t, y_t):         # :161-169
        t_x, x = x_t[0:1], x_t[1:]
        f = lambda x: self.div_y_kappa(torch.cat((x, t_x)), y_t)
        return self.laplacian_op(f)(x)

    def laplacian_x_t_laplacian_y_t_kappa(self, x_t, y_t): # :171-179
        t_x, x = x_t[0:1], x_t[1:]
        f = lambda x: self.laplacian_y_t_kappa(torch.cat((x, t_x)), y_t)
        return self.laplacian_op(f)(x)

    # name -> (rowop, colop) of the oracle's closed forms
    TABLE = {
        "kappa": ("id", "id"), "dt_y_t_kappa": ("id", "dt"), "div_y_kappa": ("id", "div"),
        "laplacian_y_t_kappa": ("id", "lap"), "dt_x_t_kappa": ("dt", "id"), "div_x_kappa": ("div", "id"),
        "laplacian_x_t_kappa": ("lap", "id"), "dt_x_t_dt_y_t_kappa": ("dt", "dt"),
        "dt_x_t_div_y_kappa": ("dt", "div"), "dt_x_t_laplacian_y_t_kappa": ("dt", "lap"),
        "div_x_dt_y_t_kappa": ("div", "dt"), "div_x_div_y_kappa": ("div", "div"),
        "div_x_laplacian_y_t_kappa": ("div", "lap"), "laplacian_x_t_dt_y_t_kappa": ("lap", "dt"),
        "laplacian_x_t_div_y_kappa": ("lap", "div"), "laplacian_x_t_laplacian_y_t_kappa": ("lap", "lap"),
    }

    # models/GP.py:630-651 + 658-663 without casts: u_hat(x) = row(x) . alpha
    def solution_function(self, x_t, x_dom, x_bdy, alpha):
        N, Nb = len(x_dom), len(x_bdy)
        parts = [torch.stack([self.kappa(x_t, y) for y in x_dom]),
                 torch.stack([self.kappa(x_t, y) for y in x_bdy]),
                 torch.stack([self.laplacian_y_t_kappa(x_t, y) for y in x_dom]),
                 torch.stack([self.dt_y_t_kappa(x_t, y) for y in x_dom]),
                 torch.stack([self.div_y_kappa(x_t, y) for y in x_dom])]
        return torch.dot(torch.cat(parts), alpha)
