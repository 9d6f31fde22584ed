"""Host mirror of halo2_proofs::poly::EvaluationDomain (poly/domain.rs) for the basis changes of the prover, each one a
single fused NTT call on the device (row a3 of SURVEY.md 8a):
  lagrange_to_coeff   = iNTT(omega^-1) with the 1/n divisor folded into the twiddle table
  coeff_to_extended   = scale coefficient i by ZETA^(i mod 3) (fused into the first load), zero-pad to 2^extended_k, NTT(omega_ext)
  extended_to_coeff   = iNTT(omega_ext^-1) * 1/N, unscale by ZETA^-(i mod 3) (fused into the last store), truncate to n*(d-1)
Constants follow EvaluationDomain::new(j = cs.degree(), k); they are derived with the device field kernels."""
import numpy as np

from . import arithmetic as A
from .params import fr_scalar_dev


class EvaluationDomain:
    def __init__(self, j, k):
        self.k, self.n = k, 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k, self.extended_n = ek, 1 << ek
        self.omega, self.omega_inv = A.root_of_unity(k)
        self.extended_omega, self.extended_omega_inv = A.root_of_unity(ek)
        inv = lambda v: A.field_unop_dev(A.FR, A.UOP_INV, fr_scalar_dev(v)).cpu().numpy().view(np.uint64)[0]
        self.ifft_divisor = inv(self.n)
        self.extended_ifft_divisor = inv(self.extended_n)

    def lagrange_to_coeff(self, values_dev):
        """(n,4) device tensor of Lagrange values -> coefficients (new tensor)."""
        out = values_dev.clone()
        return A.best_fft_dev(out, self.omega_inv, self.k, scale=self.ifft_divisor)

    def coeff_to_lagrange(self, coeffs_dev):
        return A.best_fft_dev(coeffs_dev.clone(), self.omega, self.k)

    def coeff_to_extended(self, coeffs_dev):
        import torch
        ext = torch.zeros((self.extended_n, 4), dtype=torch.int64, device=coeffs_dev.device)
        ext[: self.n] = coeffs_dev
        return A.best_fft_dev(ext, self.extended_omega, self.extended_k, coset_zeta=1)

    def extended_to_coeff(self, ext_dev):
        out = A.best_fft_dev(ext_dev.clone(), self.extended_omega_inv, self.extended_k, scale=self.extended_ifft_divisor, coset_zeta=2)
        return out[: self.n * self.quotient_poly_degree]
