
ìA((Shanghai_SH0007_2020:3,Hangzhou_ZJU-07_2020:5,Wuhan_IPBCAMS-WH-03_2019:1,Wuhan_WH01_2019:2,Wuhan_WIV07_2019:1,Shanghai_SH0093_2020:2,Wuhan_HBCDC-HB-01_2019:1,Wuhan_IPBCAMS-WH-01_2019:3,Wuhan_IVDC-HB-04_2020:2,Wuhan_IVDC-HB-05_2019:2,Wuhan_WIV05_2019:2,Sweden_01_2020:7,Taiwan_2_2020:3,USA_CA2_2020:2,Shanghai_SH0040_2020:2,Singapore_7_2020:2,Australia_VIC02_2020:2,France_IDF0515_2020:3,Hangzhou_ZJU-01_2020:3,Shanghai_SH0037_2020:5,Nepal_61_2020:1,China_WH-09_2020:1,Jiangxi_IVDC-JX-002_2020:1,Zhejiang_WZ-01_2020:1,Japan_KY-V-029_2020:3,Hangzhou_ZJU-03_2020:1,Netherlands_Utrecht_12_2020:2,Singapore_1_2020:1,Taiwan_NTU02_2020:2,England_SHEF-BFD36_2020:3,France_IDF0626_2020:1,Malaysia_MKAK-CL-2020-7554_2020:2,Cambodia_0012_2020:1,(USA_CA-CDPH-UC2_2020:4,(USA_CA9_2020:0,(USA_CA-CDPH-UC3_2020:1,USA_CA-CDPH-UC4_2020:0):0):0):1,(node_3_condensed_2_leaves:0):1,(Germany_NRW-10_2020:1,Germany_NRW-09_2020:1,Poland_PL_P1_2020:1,(Germany_NRW-04_2020:0,Germany_NRW-07_2020:1,(Germany_NRW-08_2020:0,(node_34_condensed_2_leaves:0):0):2):2,(node_15_condensed_5_leaves:0):0):2,(Malaysia_MKAK-CL-2020-5045_2020:2,(Singapore_2_2020:0,(node_25_condensed_2_leaves:0):0):0):1,(Jiangsu_JS01_2020:2,Shanghai_SH0031_2020:1):1,(Australia_NSW02_2020:1,(Hangzhou_ZJU-06_2020:1,(Shanghai_SH0085_2020:4,Shanghai_SH0094_2020:0):2):1):1,(India_1-27_2020:3,Shanghai_SH0011_2020:1,Fujian_13_2020:1,(Singapore_11_2020:1,USA_WI1_2020:0):0):1,(Wuhan_HBCDC-HB-05_2020:1,Shandong_IVDC-SD-001_2020:3,England_200960041_2020:2,Australia_NSW05_2020:1,(Georgia_Tb-82_2020:2,Australia_NSW09_2020:3,(Canada_BC_25211_2020:4,Kuwait_KU09_2020:2):1,(Kuwait_KU18_2020:1,Kuwait_KU12_2020:2,Australia_NSW06_2020:1):0):2,(node_16_condensed_2_leaves:0):2,(Taiwan_CGMH-CGU-04_2020:3,(node_26_condensed_2_leaves:0):0):3,Kuwait_KU17_2020:1,(Canada_BC_17397_2020:2,Canada_BC_13297_2020:2):1,(Australia_QLD09_2020:1,Australia_NSW07_2020:0):1,(USA_NY1-PV08001_2020:3,Canada_BC_37_0-2_2020:0,Canada_BC_69243_2020:2,(England_200990002_2020:2,Germany_BavPat2_2020:1):2):1):4,(node_4_condensed_3_leaves:0):2,(Jiangsu_JS02_2020:4,Shanghai_SH0008_2020:3,USA_CA5_2020:1):1,(Wuhan_WIV02_2019:1,(node_17_condensed_2_leaves:0):0):1,(node_5_condensed_2_leaves:0):1,(Chongqing_IVDC-CQ-001_2020:0,(Japan_Hu_DP_Kng_19-027_2020:1,(node_27_condensed_2_leaves:0):0):0):1,(node_6_condensed_2_leaves:0):0,(Hangzhou_ZJU-09_2020:0,Australia_VIC01_2020:3):1,(node_7_condensed_2_leaves:0):0,(USA_MA1_2020:3,(Shanghai_SH0058_2020:4,((node_35_condensed_2_leaves:0):1,(node_36_condensed_2_leaves:0):0):0):1):1,((France_B2334_2020:1,France_B2340_2020:1):2,((Georgia_Tb-468_2020:1,Georgia_Tb-537_2020:1,Georgia_Tb-54_2020:0):1,Italy_SPL1_2020:0):0):2,(node_8_condensed_2_leaves:0):0,(node_9_condensed_2_leaves:0):0,(England_20102000106_2020:2,(England_09c_2020:0,(Brazil_ES-225_2020:1,England_20100001406_2020:1,(England_200940527_2020:0,(node_52_condensed_2_leaves:0):0):1):0):2,(England_200960515_2020:1,(Finland_FIN01032020_2020:0,(node_37_condensed_2_leaves:0):0):1):0,(Brazil_SPBR-10_2020:1,Switzerland_1000477102_2020:2,Brazil_SPBR-02_2020:1,England_200990006_2020:0,Netherlands_Utrecht_18_2020:2):1):3,(node_10_condensed_2_leaves:0):0,(Netherlands_Utrecht_19_2020:3,Netherlands_Limburg_6_2020:2,Netherlands_Naarden_1364774_2020:0,(node_18_condensed_2_leaves:0):1,(Netherlands_Oss_1363500_2020:0,Netherlands_Tilburg_1363354_2020:2):1):1,(node_11_condensed_2_leaves:0):0,(Singapore_8_2020:1,(node_19_condensed_3_leaves:0):1):1,(Shanghai_SH0086_2020:3,(Finland_FIN-114_2020:0,(node_28_condensed_2_leaves:0):0):0,(Finland_FIN-25_2020:6,Belgium_DB-03023_2020:4,Denmark_SSI-102_2020:3,Denmark_SSI-02_2020:1,Belgium_BM-03012_2020:3,England_20100022706_2020:2,Denmark_SSI-04_2020:3,Taiwan_NTU03_2020:2,Finland_FIN03032020A_2020:2,Switzerland_GE4984_2020:2,Switzerland_AG0361_2020:2,Finland_FIN-266_2020:1,Luxembourg_Lux1_2020:1,Japan_SMU-0311S3_2020:1,Denmark_SSI-03_2020:1,Netherlands_Gelderland_2_2020:1,Georgia_Tb-273_2020:1,Belgium_SH-03014_2020:1,(((node_53_condensed_2_leaves:0):0,(node_54_condensed_2_leaves:0):0):0,(node_38_condensed_2_leaves:0):2):1,(France_HF2174_2020:2,France_GE1977_2020:3,France_HF1684_2020:0,(node_39_condensed_2_leaves:0):1):2,Belgium_VAG-03013_2020:4,(Netherlands_Utrecht_1363564_2020:1,(node_40_condensed_2_leaves:0):0):4,(England_20099107406_2020:1,Georgia_Tb_2020:0):1,(node_29_condensed_2_leaves:0):0,(node_30_condensed_2_leaves:0):0,(France_BFC2094_2020:1,France_BFC2147_2020:1,((node_55_condensed_2_leaves:0):0,(node_56_condensed_2_leaves:0):0):0):1,(Hungary_mbl1_2020:1,node_31_condensed_5_leaves:0):1,(Nigeria_Lagos01_2020:1,Netherlands_Delft_1363424_2020:1,Denmark_SSI-01_2020:1,England_200990723_2020:1,Switzerland_TI2045_2020:1,Switzerland_1000477377_2020:2,Denmark_SSI-05_2020:4,Chile_Santiago-2_2020:1,Brazil_SPBR-14_2020:1,Finland_FIN-318_2020:2,Netherlands_Limburg_5_2020:2,Belgium_DBA-03032_2020:2,Switzerland_SZ1417_2020:2,Belgium_GMH-03022_2020:2,Netherlands_Utrecht_14_2020:2,Netherlands_Utrecht_1_2020:2,Finland_FIN03032020C_2020:1,Netherlands_Berlicum_1363564_2020:1,(node_41_condensed_2_leaves:0):2,(node_42_condensed_2_leaves:0):1,(node_43_condensed_2_leaves:0):1,(Switzerland_1000477797_2020:1,Brazil_BA-312_2020:1,(Switzerland_GE8102_2020:0,(node_67_condensed_2_leaves:0):0):0):1,(Finland_FIN-313_2020:1,Netherlands_Utrecht_15_2020:0):2,((node_57_condensed_2_leaves:0):0,(node_58_condensed_2_leaves:0):0,node_44_condensed_3_leaves:0):0,(Switzerland_GE1422_2020:1,Netherlands_Diemen_1363454_2020:3,(Switzerland_GE1402_2020:0,(node_68_condensed_2_leaves:0):0):0):1,(Switzerland_BE6651_2020:1,Netherlands_NoordHolland_2_2020:1,(Switzerland_GE3895_2020:0,(node_69_condensed_5_leaves:0):0):0):1,Belgium_BC-03016_2020:2,(Switzerland_1000477757_2020:2,Brazil_SPBR-12_2020:1,Germany_Baden-Wuerttemberg-1_2020:0):1,((node_59_condensed_2_leaves:0):0,(node_60_condensed_4_leaves:0):0,(node_61_condensed_2_leaves:0):0,node_45_condensed_8_leaves:0):1,(Belgium_UMF-03025_2020:0,(node_62_condensed_2_leaves:0):1):1):3,(Netherlands_NoordHolland_3_2020:2,Netherlands_Flevoland_1_2020:0):1,(NetherlandsL_Houten_1363498_2020:1,Italy_UniSR1_2020:1,France_N1620_2020:1,(Switzerland_TI9486_2020:0,(node_63_condensed_2_leaves:0):0):0):1,(Georgia_Tb-673_2020:2,France_B2330_2020:1,France_PL1643_2020:2,France_HF1988_2020:1,France_B2348_2020:1,(France_B2351_2020:1,France_B2344_2020:1,France_B2349_2020:0):1,(France_IDF2256_2020:1,France_HF1871_2020:2):1,((node_64_condensed_2_leaves:0):0,(node_65_condensed_2_leaves:0):0,node_46_condensed_2_leaves:0):0,Netherlands_Haarlem_1363688_2020:1):2,(Ireland_Limerick-19933_2020:0,(node_47_condensed_2_leaves:0):0,(node_48_condensed_2_leaves:0):0):0,node_20_condensed_6_leaves:0):1):3,node_1_condensed_15_leaves:0):0,(Wuhan_HBCDC-HB-06_2020:3,Wuhan_HBCDC-HB-02_2020:2,Wuhan_HBCDC-HB-04_2020:3,Shanghai_SH0075_2020:2,Shanghai_SH0010_2020:2,Tianmen_HBCDC-HB-07_2020:5,Shanghai_SH0041_2020:4,India_1-31_2020:4,Shanghai_SH0009_2020:1,Malaysia_MKAK-CL-2020-5096_2020:2,(France_GE1583_2020:1,Georgia_Tb-390_2020:1,Spain_Valencia6_2020:1,Spain_CastillayLeon201437_2020:1,((node_32_condensed_2_leaves:0):0,node_21_condensed_4_leaves:0):0):5,(Sichuan_IVDC-SC-001_2020:1,USA_CA1_2020:1,USA_IL2_2020:2,(Vietnam_CM99_2020:2,(Shanghai_SH0032_2020:0,(node_49_condensed_2_leaves:0):0):0):0):3,(Shanghai_SH0035_2020:2,Australia_QLD02_2020:2,(Australia_QLD03_2020:1,(Australia_VIC07_2020:1,((Netherlands_Utrecht_16_2020:1,Netherlands_Utrecht_17_2020:0):1,(node_66_condensed_2_leaves:0):0,node_50_condensed_4_leaves:0):0):0):0):2,(Shandong_LY003_2020:2,Shandong_LY005_2020:2,(Shandong_LY006_2020:0,(node_33_condensed_2_leaves:0):0):0):1,(node_12_condensed_2_leaves:0):4,England_02_2020:3,(Chongqing_YC01_2020:2,(((USA_CA-CDPH-UC7_2020:0,USA_CA-CDPH-UC9_2020:1):3,(node_51_condensed_2_leaves:0):2):2,(USA_WA1_2020:0,(Fujian_8_2020:0,(Malaysia_MKAK-CL-2020-6430_2020:1,Hangzhou_ZJU-08_2020:0):0):0):0):0):1,(Shanghai_SH0059_2020:3,USA_TX1_2020:4,(Shanghai_SH0004_2020:1,(Germany_BavPat3_2020:1,Shanghai_SH0005_2020:0):0):0,(Japan_TY-WK-012_2020:1,node_22_condensed_2_leaves:0):1):1,(Shanghai_SH0024_2020:0,(node_23_condensed_2_leaves:0):0):0,(Yunnan_IVDC-YN-003_2020:0,USA_AZ1_2020:1,Anhui_SZ005_2020:6):1,(Beijing_233_2020:1,(node_24_condensed_3_leaves:0):0):2,(node_13_condensed_2_leaves:0):0,(node_14_condensed_2_leaves:0):0,node_2_condensed_3_leaves:0):2):0;  #
ß±"
≤æ"
¢◊";

˝
"

 L"
Èõ"
≥≈"
ﬂ◊"

‘6"

∏6" 

Ù[" 
¡>"
Íÿ"
≥·"
ò©" 
¬"
ƒA"

õF" 
Â÷"
›‹"
æ°" 
«°" 

Ë6" 
ë•"O

ù" 
∫H"

©g"

™g"
‡á"
êª"
†Ã""

º~"
Ï "
†Ã"
ËÑ"
†Ã"
ºâ" 
†Ã"
†Ã"
±·"

Ñ+"
†Ã"$

¶	"

ﬁI"

ÀV"&

ÀV"
∑∏" 
†Ã"6
Ω"
ºC"

ÀV"
õ•"
†Ã"
‚ª"

∫"

“@" 

«"%

•Z"

‹w"
˜‰"
ŒÉ"

§^"

Õb"	
‰√"
 F"

ìJ""

ıS"
Æû"
ÛÁ"
¡ƒ"

ﬁ"

ÏR"
∂»"

ƒM"+
ôz"
àÈ"
âÈ"
äÈ"   

ﬂM" 
Î≈" 

†" 

À" 

ˆ<"
…Ÿ"

∆" 

Ø$"
≥Æ" 

∂i"

ÊA" 
ÍÊ"      
ã‘"

é"
˚”"    

∏"
ø "
§Ë"	
⁄Ä"
âÊ"

áA"

≠" 

÷g"

Ò2"
Ôû" 2

Ê0" 

œx" 
‚Ö"
∞®" 
›á"%

Â"

¡r"
Å≤"

‡d"
ªË"  
Ôj" 2

ı
" 

ÀV"
ê‡"
ÆË"

Ω[" %

É9"

•T"
‹◊"

Ÿ_"
ö©"	
œú"

Ù"

ÕC"

≠N"

ªV"'
Ωæ"
©ﬁ" 
ªÁ"	
¸ "0

°"

∑<"

¬V"

‚n"
ø"
€ú" 

§l"
ê"
›◊"
»¬"
”P"

¬V" &

£"
˛∫"
Ëﬂ"!
Ω'"

¬V"
«¶"   

Ωr" 
úò"
˘"
˘∏"

π"
éû"
æÂ" 

¬V" 
™J"&

™" 
˛ƒ"
„‚" 
úò"
Ø†"

áJ"

ÀM"
ø"
ïF"

∞E"
Ö±"
†Ã" 
À©"1

î" 

¸9"

Éj"
„Ã"%

À"

©t"
•œ"

´"	
Öæ"
ƒ¶"   
†Ã" 

ÀV"  
√Á"    
üÆ" '
˘î"
†Ã"
øË"   
∂·""

æ"
èà"
‚ª"
‹'"0

Ô""

Ù#"

Ë7" 

‡D" 

Ê"    

ÀV"
†Ã"

ì"
ó£" 

Ì)"
¨¥" 
ì"

œP"
‚∑"      %

ÀV"

’s"
†Ã"
õπ"
¯’"
∞"

˛"  

«""
Ú“"
–§"    

ıS"
›Â"   
ﬂÜ"

ÀV"
∫æ"
é“"

‘" 
îd"

øf"  

Ç"&

à+"
üú"
ÿ„"

À" 
∞∫" 
Çà" 
¬" 

∆="
Çà"  
†Ã"
„™"

öO" !

Ò"

›"
Î∂" 

¿	"

À/"
Õ~"    

»p"G
â"

ù"

≤" 
·´" 
§Æ"
ÀÃ"-

∫"

»"

—"
œÆ"'
·∂"
ìÈ" 
ñÈ" 

⁄h""

⁄h"
≈æ"
∆æ"

ø6"

¡6" !
ª"

⁄h"
õ£"
Ωì"
€«"
—≤"
ëÁ"

¶l"
§·"
Æü" 
¢œ"	
¢œ"
ó∏"
ØË"

‡9"	
û¬"
ê‡"
à?"
Ü∆"     
“ü"
Í—" 

"
€«"
…c"
äû"&

“"
∫æ"
ÁÊ" 

≤c" 3

⁄h"
¥ù"
∆æ"
ö·"-

ﬂ"

ú2"

ÀV"
Á«"

Ì "   
çº"
«®"     

‹w"

ê"

˚e"     	
¨û"
ã«"  '
—·" 
“·" 
”·"

Ω" 
¬"	
Ã¥"

¬\" 

≈`"

≠" 
ÓÂ"1

¢"

ÒN" 

‡i"
≥π"

¡"
Äƒ"

˝"
¶”"

ŸF"
¶”"
„ñ"
¶”"

ÀV"
¶”"

⁄h"
¶”"

ƒ|"
¶”"

Øx"
¶”"
¶”"
¶”"
ÆY"
¶”" 

¯f" 

øi" 

π"	
ò≠"	
¡“"    
ÖÆ" 
¶”"
”∏"       
‹„"

“T"!

∫"
¬"
èº"     
ˇö"

Õ5"

…"    
≈æ"
∆æ"

ôP" 

ø"
‚◊"

ÀV" 
¶”"       
˚l" 

⁄h" 

£"
©5"

Æh" 
ª"
¬"
¨6"

ËS"    

£"
€«"

È
"

ˆB"

¬d" 

ë^"

—i"	
û¢"
Ë+"
†´" 
ì"
ã" 

÷R"
õπ"

”"
≥”"      

Ç"        

ŒD"
€"

ô," 
Î∂"
˝»"

Ì^" 
à∫"'
¥î"
≤æ"
Â‚"
‚◊"
¢‡"

Ë"
ﬂ™"9
Á"

Î" 

÷p"

˘q"
õ›",

œ"

≤"

€1"
¶Y".
õ"

Â2"
ÌÉ"
üæ"
‹á"

¶""

’,"?

ÖJ" 

’s"
˚ "
Òﬂ"
ø·"
—«"

ˇw" 

™)"
ñë"    '
‚ª"
È–"
≠€"
ˆï" 

å" 

Í" 

È" 
´H"
ñ•"    
Œ·" 
ÆË" 
˛´"
†‚" 
û√"
Û«" 
´ë" 

∂N" 

˝" 

Ô0"    

„"
˜õ"
€ﬂ" 
˘Ë" 
˚Ë"    4
æà"
€ì"
ËÀ"
§ﬂ" #
∏ê"
µ∏"
úÁ"
åç"

æ
"
ê‰" 
”ä"
¬ã"
Ê"
”Ä"
ëµ" 
œÑ"

ˆ<"
¥ù"      

â/" 
ß„"&

ù"
≤ª"
€ "0
´ë"
üî" 
Áï"
ï⁄" 
≥∂" 
‹∫" 

Ê"

–"     

ÀV" 
ß„"E

ç"

é"
Ê'"

ÀV"

±}"
‚Ö" 

≤""

∆'"	
ı‰"       G
node_3_condensed_2_leavesAustralia_NSW08_2020Australia_NSW10_2020M
node_34_condensed_2_leavesGermany_NRW-01_2020Netherlands_Limburg_4_2020á
node_15_condensed_5_leavesGermany_NRW-05_2020Germany_NRW-03_2020Germany_NRW-02-1_2020Germany_NRW-06_2020Brazil_SPBR-11_2020O
node_25_condensed_2_leavesSingapore_6_2020Malaysia_MKAK-CL-2020-5047_2020H
node_16_condensed_2_leavesShanghai_SH0022_2020Shanghai_SH0023_2020N
node_26_condensed_2_leavesTaiwan_CGMH-CGU-03_2020Taiwan_CGMH-CGU-05_2020`
node_4_condensed_3_leavesFrance_IDF0373_2020France_IDF0386-islP1_2020France_IDF0372_2020B
node_17_condensed_2_leavesUSA_CA8_2020Wuhan_HBCDC-HB-02_2019G
node_5_condensed_2_leavesAustralia_NSW03_2020Australia_VIC03_2020K
node_27_condensed_2_leavesSingapore_3_2020Japan_Hu_DP_Kng_19-020_2020K
node_6_condensed_2_leavesJapan_NA-20-05-1_2020Taiwan_CGMH-CGU-01_2020G
node_7_condensed_2_leavesJapan_OS-20-07-1_2020Chongqing_ZX01_20208
node_35_condensed_2_leavesUSA_CA3_2020USA_CA4_2020L
node_36_condensed_2_leavesCanada_ON-PHL2445_2020Canada_ON-VIDO-01_2020N
node_8_condensed_2_leavesJiangsu_IVDC-JS-001_2020Hangzhou_HZCDC0001_2020D
node_9_condensed_2_leavesHangzhou_ZJU-04_2020Jiangsu_JS03_2020L
node_52_condensed_2_leavesEngland_200990725_2020England_200990724_2020M
node_37_condensed_2_leavesFinland_FIN-274_2020Finland_FIN03032020B_20209
node_10_condensed_2_leavesUSA_CA6_2020Taiwan_4_2020c
node_18_condensed_2_leavesNetherlands_Dalen_1363624_2020%Netherlands_Loon_op_zand_1363512_2020G
node_11_condensed_2_leavesFrance_RA739_2020England_200690300_2020S
node_19_condensed_3_leavesSingapore_5_2020Singapore_9_2020Singapore_10_2020H
node_28_condensed_2_leavesGermany_BavPat1_2020Shanghai_SH0014_2020P
node_53_condensed_2_leavesEngland_20100004706_2020England_20100004806_2020N
node_54_condensed_2_leavesEngland_20100005406_2020England_200990660_2020P
node_38_condensed_2_leavesEngland_20100121006_2020England_20100121007_2020D
node_39_condensed_2_leavesFrance_HF1995_2020France_HF2239_2020Z
node_40_condensed_2_leavesNetherlands_Utrecht_6_2020 Netherlands_Utrecht_1363628_2020F
node_29_condensed_2_leavesBrazil_SPBR-06_2020Brazil_SPBR-05_2020E
node_30_condensed_2_leavesBrazil_SPBR-09_2020France_GE1973_2020I
node_55_condensed_2_leavesFrance_HF2196_2020Switzerland_GE6679_2020N
node_56_condensed_2_leavesSwitzerland_BE2536_2020Switzerland_GE4135_2020ë
node_31_condensed_5_leavesSwitzerland_GE5373_2020Switzerland_GE3121_2020Panama_328677_2020Georgia_Tb-712_2020Spain_Galicia201663_2020T
node_41_condensed_2_leavesNetherlands_Limburg_3_2020Netherlands_Utrecht_2_2020N
node_42_condensed_2_leavesSwitzerland_GR2988_2020Switzerland_GR3043_2020N
node_43_condensed_2_leavesSwitzerland_VD0503_2020Switzerland_VD5615_2020F
node_67_condensed_2_leavesBrazil_SPBR-08_2020Brazil_SPBR-13_2020S
node_57_condensed_2_leavesSwitzerland_1000477796_2020England_20099038206_2020D
node_58_condensed_2_leavesVietnam_CM296_2020Vietnam_CM295_2020j
node_44_condensed_3_leavesMexico_CDMX-InDRE_01_2020Vietnam_39607_2020Netherlands_Overijssel_2_2020W
node_68_condensed_2_leavesSwitzerland_AG7120_2020 Netherlands_Helmond_1363548_2020ù
node_69_condensed_5_leavesSwitzerland_GE0199_2020Switzerland_BL0902_2020Switzerland_GE9586_2020Switzerland_1000477806_2020Switzerland_BS0914_2020T
node_59_condensed_2_leavesNetherlands_Utrecht_4_2020Netherlands_Limburg_2_2020á
node_60_condensed_4_leavesBrazil_SPBR-03_2020Netherlands_Utrecht_5_2020Netherlands_Gelderland_1_2020Ireland_Dublin-19072_2020W
node_61_condensed_2_leavesNetherlands_Utrecht_7_2020Netherlands_Overijssel_1_2020Ï
node_45_condensed_8_leavesNetherlands_NoordHolland_1_2020Netherlands_Utrecht_10_2020Germany_NRW-011_2020Netherlands_Utrecht_13_2020Netherlands_Gelderland_3_2020Finland_FIN-508_2020Brazil_SPBR-04_2020Brazil_SPBR-07_2020L
node_62_condensed_2_leavesBelgium_DBD-03024_2020Belgium_QKJ-03015_2020K
node_63_condensed_2_leavesFrance_N2223_2020Netherlands_Utrecht_3_2020B
node_64_condensed_2_leavesFrance_B2336_2020France_B2335_2020B
node_65_condensed_2_leavesFrance_B2343_2020France_B2346_2020E
node_46_condensed_2_leavesFrance_B2337_2020Finland_FIN-455_2020N
node_47_condensed_2_leavesDenmark_SSI-09_2020Ireland_Limerick-19935_2020U
node_48_condensed_2_leavesNetherlands_Utrecht_11_2020Netherlands_Utrecht_8_2020Æ
node_20_condensed_6_leavesItaly_CDG1_2020!Netherlands_Zeewolde_1365080_2020Georgia_Tb-477_2020Japan_SMU-0311S2_2020Brazil_SPBR-01_2020Ireland_Limerick-19934_2020€
node_1_condensed_15_leavesWuhan_HBCDC-HB-03_2019Wuhan_IVDC-HB-01_2019Wuhan_WIV06_2019Wuhan_WIV04_2019Wuhan_IPBCAMS-WH-04_2019Wuhan_IPBCAMS-WH-02_2019Wuhan_WH03_2020Zhejiang_WZ-02_2020Nonthaburi_61_2020Nonthaburi_74_2020Hangzhou_ZJU-05_2020Hangzhou_HZ-1_2020Finland_1_2020England_200641094_2020England_200690756_2020R
node_32_condensed_2_leavesChile_Santiago_op3d1_2020Chile_Santiago_op4d1_2020z
node_21_condensed_4_leavesSpain_Valencia4_2020Chile_Santiago_op2d1_2020Spain_Valencia5_2020Chile_Santiago-1_2020I
node_49_condensed_2_leavesVietnam_VR03-38142_2020Vietnam_38142_2020D
node_66_condensed_2_leavesAustralia_QLD04_2020Singapore_4_2020l
node_50_condensed_4_leavesAustralia_QLD01_2020Shanghai_SH0002_2020Shanghai_SH0003_2020USA_CA7_2020F
node_33_condensed_2_leavesShandong_LY004_2020Shandong_LY008_2020D
node_12_condensed_2_leavesChile_Talca-1_2020Chile_Talca-2_2020H
node_51_condensed_2_leavesUSA_CA-CDPH-UC5_2020USA_CA-CDPH-UC6_2020H
node_22_condensed_2_leavesJapan_TY-WK-501_2020Japan_TY-WK-521_2020H
node_23_condensed_2_leavesShanghai_SH0043_2020Shanghai_SH0013_2020R
node_24_condensed_3_leavesBeijing_235_2020Beijing_231_2020Beijing_105_2020A
node_13_condensed_2_leavesAustralia_NSW01_2020Taiwan_3_2020G
node_14_condensed_2_leavesTaiwan_NTU01_2020Belgium_GHB-03021_2020Z
node_2_condensed_3_leavesWuhan_WH04_2020Wuhan_HBCDC-HB-03_2020Hangzhou_ZJU-02_2020" " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " " 