"""MarginLoss (OpenKE/openke/module/loss/MarginLoss.py:10-32; the paper's copy module/loss.py:5-28).  The plain form
is fused into mre_transe_margin_step by the NegativeSampling strategy; this module is the general (torch, [B, neg]
elementwise) form used when a caller combines it with other scorers or the self-adversarial weights."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class MarginLoss(nn.Module):
    def __init__(self, adv_temperature=None, margin=6.0):
        super().__init__()
        self.margin = nn.Parameter(torch.Tensor([margin]))
        self.margin.requires_grad = False
        if adv_temperature is not None:
            self.adv_temperature = nn.Parameter(torch.Tensor([adv_temperature]))
            self.adv_temperature.requires_grad = False
            self.adv_flag = True
        else:
            self.adv_flag = False

    def get_weights(self, n_score):
        return F.softmax(-n_score * self.adv_temperature, dim=-1).detach()

    def forward(self, p_score, n_score):
        if self.adv_flag:
            return (self.get_weights(n_score) * torch.max(p_score - n_score, -self.margin)).sum(dim=-1).mean() + self.margin
        return (torch.max(p_score - n_score, -self.margin)).mean() + self.margin

    def predict(self, p_score, n_score):
        score = self.forward(p_score, n_score)
        return score.cpu().data.numpy()
