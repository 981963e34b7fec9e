"""LatticeNet (`LNN`): the U-Net over lattice levels that drives the hot path in the reference's
training / evaluation scripts (/root/reference/latticenet_py/lattice/models.py:70-266).  Same module
attribute names (`point_net`, `resnet_blocks_per_down_lvl_list`, `coarsens_list`,
`resnet_blocks_bottleneck`, `finefy_list`, `resnet_blocks_per_up_lvl_list`, `slice_fast_cuda`) so a
reference `state_dict` maps one to one.
"""
import torch

from . import lattice as _lattice
from .lattice_modules import (BottleneckBlock, CoarsenAct, DistributeLatticeModule, GnReluFinefy, PointNetModule,
                              ResnetBlock, SliceFastCUDALatticeModule, filter_readings)


class LNN(torch.nn.Module):
    def __init__(self, nr_classes, model_params, device=None, verbose=False):
        super().__init__()
        mp = model_params
        self.nr_classes = nr_classes
        self.model_params = mp
        self.nr_downsamples = mp.nr_downsamples()
        self.nr_blocks_down_stage = list(mp.nr_blocks_down_stage())
        self.nr_blocks_bottleneck = mp.nr_blocks_bottleneck()
        self.nr_blocks_up_stage = list(mp.nr_blocks_up_stage())
        self.nr_levels_down_with_normal_resnet = mp.nr_levels_down_with_normal_resnet()
        self.nr_levels_up_with_normal_resnet = mp.nr_levels_up_with_normal_resnet()
        compression = mp.compression_factor()
        log = print if verbose else (lambda *a, **k: None)

        self.grad_sync_hook = None       # callable(grad) -> None, see forward(); set by graphed.GraphedTrainStep for world > 1
        self.fused_point_net = True      # False: the module-by-module path of the reference (distribute rows, torch MLP, scatter ops)
        self.distribute = DistributeLatticeModule()
        self.pointnet_channels_per_layer = list(mp.pointnet_channels_per_layer())
        self.start_nr_filters = mp.pointnet_start_nr_channels()
        self.point_net = PointNetModule(self.pointnet_channels_per_layer, self.start_nr_filters, device=device)

        # ---- encoder: blocks at each level, then coarsen (channels x2 x compression) -------------
        self.resnet_blocks_per_down_lvl_list = torch.nn.ModuleList([])
        self.coarsens_list = torch.nn.ModuleList([])
        skip_channels = []
        ch = self.start_nr_filters
        for lvl in range(self.nr_downsamples):
            blocks = torch.nn.ModuleList([])
            for _ in range(self.nr_blocks_down_stage[lvl]):
                if lvl < self.nr_levels_down_with_normal_resnet:
                    log("down resnet block", ch)
                    blocks.append(ResnetBlock(ch, ch, [1, 1], [False, False], False, device=device))
                else:
                    log("down bottleneck block", ch)
                    blocks.append(BottleneckBlock(ch, ch, [False, False, False], device=device))
            self.resnet_blocks_per_down_lvl_list.append(blocks)
            skip_channels.append(ch)
            ch_coarse = int(ch * 2 * compression)
            log("coarsen ->", ch_coarse)
            self.coarsens_list.append(CoarsenAct(ch, ch_coarse, device=device))
            ch = ch_coarse

        # ---- bottleneck ----------------------------------------------------------------------------
        self.resnet_blocks_bottleneck = torch.nn.ModuleList(
            [BottleneckBlock(ch, ch, [False, False, False], device=device) for _ in range(self.nr_blocks_bottleneck)])

        # ---- decoder: finefy (channels /2), concat the skip, blocks ---------------------------------
        self.do_concat_for_vertical_connection = True
        self.finefy_list = torch.nn.ModuleList([])
        self.resnet_blocks_per_up_lvl_list = torch.nn.ModuleList([])
        for lvl in range(self.nr_downsamples):
            skip = skip_channels.pop()
            ch_fine = int(ch / 2)
            log("finefy ->", ch_fine)
            self.finefy_list.append(GnReluFinefy(ch, ch_fine, device=device))
            ch = skip + ch_fine if self.do_concat_for_vertical_connection else skip
            blocks = torch.nn.ModuleList([])
            for j in range(self.nr_blocks_up_stage[lvl]):
                # the very last conv feeds the slice (no norm after it), so it carries a bias
                last = j == self.nr_blocks_up_stage[lvl] - 1 and lvl == self.nr_downsamples - 1
                if lvl >= self.nr_downsamples - self.nr_levels_up_with_normal_resnet:
                    log("up resnet block", ch)
                    blocks.append(ResnetBlock(ch, ch, [1, 1], [False, last], False, device=device))
                else:
                    log("up bottleneck block", ch)
                    blocks.append(BottleneckBlock(ch, ch, [False, False, last], device=device))
            self.resnet_blocks_per_up_lvl_list.append(blocks)

        self.slice_fast_cuda = SliceFastCUDALatticeModule(in_channels=ch, nr_classes=nr_classes,
                                                          dropout_prob=mp.dropout_last_layer(), experiment="none", device=device)
        self.logsoftmax = torch.nn.LogSoftmax(dim=1)

    def forward(self, ls, positions, values):
        if positions.is_cuda:
            # tensor-core slabs of every filter bank / 1x1 weight, forward and transposed readings, in one launch
            # (nothing is launched while the weights have not changed since the last call)
            _lattice.prepare_filters(filter_readings(self, torch.is_grad_enabled()))
        if positions.is_cuda and self.fused_point_net and self.point_net.fused_supported(positions.shape[1], values.shape[1]):
            # distribute + PointNet without the [N(d+1) x ...] row tensors (csrc/ln_pointnet.cu)
            with torch.no_grad():
                ls.begin_splat(True)
                ls, indices, weights = ls.distribute_structure(positions, values, True)
            self.last_level1_lattice = ls
            self.last_level_lattices = [ls]
            lv, ls = self.point_net.forward_fused(ls, positions, values, indices, weights)
        else:
            with torch.no_grad():
                ls, distributed, indices, weights = self.distribute(ls, positions, values)
            self.last_level1_lattice = ls          # kept for inspection / tests (vertex numbering of this pass)
            self.last_level_lattices = [ls]        # one handle per lattice level of this pass (level 1 first)
            lv, ls = self.point_net(ls, distributed, indices)

        fine_structures, fine_values = [], []
        for lvl in range(self.nr_downsamples):
            for block in self.resnet_blocks_per_down_lvl_list[lvl]:
                lv, ls = block(lv, ls)
            fine_structures.append(ls)
            fine_values.append(lv)
            lv, ls = self.coarsens_list[lvl](lv, ls)
            self.last_level_lattices.append(ls)

        for block in self.resnet_blocks_bottleneck:
            lv, ls = block(lv, ls)
        if self.grad_sync_hook is not None and lv.requires_grad:
            # fires in the backward pass once every layer after this point has its gradients (decoder + slice head: most of
            # the parameter bytes): scene-parallel training starts their all-reduce here, under the rest of the backward
            lv.register_hook(self.grad_sync_hook)

        for lvl in range(self.nr_downsamples):
            skip_values = fine_values.pop()
            fine_structure = fine_structures.pop()
            lv, ls = self.finefy_list[lvl](lv, ls, fine_structure)
            lv = torch.cat((lv, skip_values), 1) if self.do_concat_for_vertical_connection else lv + skip_values
            for block in self.resnet_blocks_per_up_lvl_list[lvl]:
                lv, ls = block(lv, ls)

        logits = self.slice_fast_cuda(lv, ls, positions, indices, weights)
        return self.logsoftmax(logits), logits

    def compute_class_weights(self, class_frequencies, background_idx):
        """Inverse-log class weights (models.py:268-280)."""
        freq = torch.as_tensor(class_frequencies, dtype=torch.float32)
        w = 1.0 / torch.log(1.05 + freq)
        w[background_idx] = 1e-8
        return w
