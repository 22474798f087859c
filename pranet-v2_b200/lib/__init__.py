"""Mirror of the reference's `binary_seg/lib` import paths:
`from lib.pranet import PraNet_V2, PVT_PraNet_V2` / `from lib.PraNet_Res2Net import PraNet, PVT_PraNet`
become `from pranet_v2_b200.lib.pranet import ...` / `from pranet_v2_b200.lib.PraNet_Res2Net import ...`."""
