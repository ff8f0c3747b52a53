# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Edward Diener 2011.                                    *
#  *     (C) Copyright Paul Mensonides 2011.                                  *
#  *     Distributed under the Boost Software License, Version 1.0. (See      *
#  *     accompanying file LICENSE_1_0.txt or copy at                         *
#  *     http://www.boost.org/LICENSE_1_0.txt)                                *
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_TUPLE_ENUM_HPP
# define BOOST_PREPROCESSOR_TUPLE_ENUM_HPP
#
# include <libint2/boost/preprocessor/tuple/rem.hpp>
#
# /* BOOST_PP_TUPLE_ENUM */
#
# define BOOST_PP_TUPLE_ENUM BOOST_PP_TUPLE_REM_CTOR
#
# endif
